#!/bin/bash
# round-end evidence on one box: parity suite, every bench workload, reference arm, 2^20 lines, ncu launch list + full capture
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; echo "verify rc=$?"
for w in verify_pk verify_rand hash_g1 hash_g2 pairing aggregate; do
  python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
for w in verify verify_rand aggregate hash_g1 hash_g2; do
  python bench.py --steps 3 --warmup 3 --workload $w --log2n 20 --cpu-sample 4096 > gpurun_out/bench_${w}_2e20.json 2> gpurun_out/bench_${w}_2e20.err; echo "$w 2^20 rc=$?"
done
bash tools/gpu_profile_round.sh
