#!/bin/bash
# Launch list of one bench run (our kernels only): plain run first, then the same command under ncu; summarised per kernel
# by tools/launch_summary.py (cold-cache, serialised timings: compare SHARES, not absolutes).
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none \
   -k regex:"tc_gemm|swin_mlp|window_attn|layernorm|score_images|conv_last|drct_head|quantize_u8" -c 4000 \
   --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu exit $?"
