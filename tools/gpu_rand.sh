#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "randomized or c5_shape or multi_device" > gpurun_out/pytest_rand.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_rand.log)"
timeout 900 python bench.py --steps 3 --warmup 3 --sub c5 --c5-log2n ${C5:-23} --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_c5.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c5.json'))
print(json.dumps(d['c5']['exact']), json.dumps(d['c5']['randomized']))
"
