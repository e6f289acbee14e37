#!/bin/bash
# slot-machine bring-up: parity tests with every shape, then the bench with the slot machine forced
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shape or slot or one-thread or lane-pair or vs_oracle or aggregate_verify or ragged or multi_miller" > gpurun_out/pytest_vm.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_vm.log
for L in 0 3; do
BN254_LANES=$L timeout 600 python bench.py --steps 10 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; tail -3 gpurun_out/bench_l$L.err
python -c "
import json
d=json.load(open('gpurun_out/bench_l$L.json'))
print('lanes $L value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['kernel'], d['clocks'])
for k,v in d['roofline']['kernels'].items(): print(' ',k,v)
"
done
