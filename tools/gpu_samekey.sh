#!/bin/bash
# the same-key randomised path: parity tests in every shape, then (optional) its bench sub-record
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "same_key" > gpurun_out/pytest_samekey.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_samekey.log)"
grep -E "Error|error|assert|FAILED" gpurun_out/pytest_samekey.log | head -30
if [ -n "$1" ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 --sub same_key --no-cpu-baseline > gpurun_out/bench_samekey.json 2> gpurun_out/bench_samekey.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_samekey.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_samekey.json'))
print(json.dumps(d.get('same_key'), indent=1))
"
fi
