#!/bin/bash
# round-end evidence: plain run, ncu launch list of the same command, one full capture of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_miller2_fixed|k_final_exp|k_hash_g1' -s 3 -c 3 -o gpurun_out/prof_r01c -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
