#!/bin/bash
mkdir -p gpurun_out
M=${M:-262144} python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1; echo "gemm_bench exit $?"; cat gpurun_out/gemm_bench.txt
if [ "${NCU:-0}" = "1" ]; then
GEMM_ONLY=1 python tools/gemm_bench.py > gpurun_out/plain2.log 2>&1 &&
GEMM_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s ${SKIP:-0} -c ${COUNT:-4} -o gpurun_out/prof_gemm -f python tools/gemm_bench.py > gpurun_out/ncu2.log 2>&1
echo "ncu exit $?"
fi
