#!/bin/bash
# Launch list of one bench step (our kernels only): plain run first, then the same command under ncu.
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none \
   -k regex:"tc_gemm|window_attn|layernorm|score_images|conv_last|drct_head|quantize_u8" -s 1467 -c 489 \
   --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu exit $?"
