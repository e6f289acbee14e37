#!/bin/bash
# ncu full capture for a variant .so: usage VAR=libname KREG=regex
mkdir -p gpurun_out
export BN254_SO=$PWD/build/variants/$VAR.so
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$KREG" -s 2 -c 2 -o gpurun_out/prof_$VAR -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
