#!/bin/bash
# bench (plain) then the ncu launch list of the same command (shares, not absolutes)
mkdir -p gpurun_out
STEPS=${STEPS:-5}
timeout 600 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${NCU:-0}" = "1" ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 420 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?"
fi
