#!/bin/bash
# one full ncu capture of the slot-machine kernels of the verify workload (after a plain run of the same command)
mkdir -p gpurun_out
export BN254_LANES=${LANES:-3}
CMD="python bench.py --steps 2 --warmup 1 --sub none --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k_vm_miller2_fixed|k_vm_final_exp}" -s 2 -c 2 -o gpurun_out/${NCU_OUT:-prof_vm} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log; tail -c 300 gpurun_out/plain.log
