#!/bin/bash
# bench lines for every workload + ncu launch list + one full ncu capture of the two top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
SMI=$!
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; echo "verify rc=$?"
for w in verify_pk hash_g1 hash_g2 pairing aggregate; do
  python bench.py --steps 3 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
kill $SMI
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_miller2_fixed|k_final_exp|k_hash_g1' -c 3 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -c 1500 gpurun_out/bench_verify.json
