#!/bin/bash
# quick check of a kernel change: the pairing / verify parity tests, then the verify bench (kernel times only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vs_oracle or gt_generator or full_size_property or aggregate_verify" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_quick.log)"
timeout 600 python bench.py --steps 10 --warmup 3 --sub none --no-cpu-baseline "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_quick.err
python -c "
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), ' '.join('%s=%.2f'%(k.split('[')[0],v['ms_per_launch']) for k,v in d['roofline']['kernels'].items()))
"
