#!/bin/bash
# Launch list of one bench run (our kernels only): plain run first, then the same command under ncu; summarised per kernel
# by tools/launch_summary.py (cold-cache, serialised timings: compare SHARES, not absolutes).  CUDA-graph replay is switched
# off for both runs so that every launch is an individual kernel node for ncu.
mkdir -p gpurun_out
TAG=${TAG:-r01c}
export ADSR_CUDA_GRAPH=0
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none \
   -k regex:"tc_gemm|swin_mlp|swin_attn|window_attn|layernorm|score_images|conv_last|drct_head|quantize_u8" -c 4000 \
   --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch list exit $?"
# one full capture of the dominant kernel: the five block shapes of one RDG (launches 5..9 = second RDG of the first warm-up step)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_attn.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swin_attn_kernel -s 5 -c 5 -o gpurun_out/prof_swin_attn -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "ncu full exit $?"
