#!/bin/bash
# Runs each GPU test file in its own process (a hung kernel must not hide the other results).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in elementwise scoring attention swin_attn gemm mlp drct drn metrics_api evaluate; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 600 --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== $f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$f.log | tee -a gpurun_out/summary.txt
done
