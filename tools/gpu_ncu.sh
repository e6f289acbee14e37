#!/bin/bash
# one full ncu capture of the top kernels of the verify step (after a plain run of the same command exited 0).
# Read back here with tools/ncu_summary.py and tools/ncu_traffic.py (-> profiles/traffic.json).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --sub none --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k_miller2_fixed|k_final_exp|k_hash_g1}" -s 3 -c 3 -o gpurun_out/${NCU_OUT:-prof} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log; tail -c 300 gpurun_out/plain.log
