#!/bin/bash
# ncu launch list of the verify step (after a plain run of the same command exited 0): kernel shares of the step
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --sub none --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -c 300 gpurun_out/plain.log
