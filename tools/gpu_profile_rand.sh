#!/bin/bash
# randomised batch verification: plain run, then ncu launch list and one full capture of its kernels (2^18 tuples)
mkdir -p gpurun_out
CMD="python bench.py --workload verify_rand --log2n 18 --steps 2 --warmup 1 --no-cpu-baseline"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_verify.json 2> gpurun_out/e2e_verify.err; echo "verify rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload verify_rand > gpurun_out/e2e_verify_rand.json 2> gpurun_out/e2e_verify_rand.err; echo "verify_rand rc=$?"
$CMD > gpurun_out/plain_rand.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rand.csv $CMD > gpurun_out/ncu_launches_rand.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_rand2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_miller_twin|k_rbv_scale_g1$|k_fp12_seg_reduce' -s 3 -c 4 -o gpurun_out/prof_r01_rand -f $CMD > gpurun_out/ncu_full_rand.log 2>&1
echo "full rc=$?"
