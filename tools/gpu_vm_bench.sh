#!/bin/bash
mkdir -p gpurun_out
for L in ${LANES_LIST:-3}; do
BN254_LANES=$L timeout 600 python bench.py --steps 5 --warmup 3 --sub none --no-cpu-baseline "$@" > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; tail -3 gpurun_out/bench_l$L.err
python -c "
import json
d=json.load(open('gpurun_out/bench_l$L.json'))
print('lanes $L value',d['value'],'e2e',d['e2e']['value'])
for k,v in d['roofline']['kernels'].items(): print(' ',k,v['ms_per_launch'])
"
done
