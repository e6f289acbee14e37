#!/bin/bash
# round-end evidence on one box, no profiler: smoke, the GPU parity suite, the one-GPU bench line with every sub-record, the CPU arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['kernel'], d['clocks'])
for k, v in d['roofline']['kernels'].items(): print(' ', k, v)
for k in ('c1', 'c3', 'pairing', 'prepared', 'c4', 'c5'): print(k, json.dumps(d.get(k)))
print('cpu', d['cpu_baseline'])
print('reference', open('gpurun_out/bench_reference.json').read()[:400])
PY
