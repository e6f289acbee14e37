#!/bin/bash
# one-GPU bench run (all sub-records) + the reference arm; args are passed to bench.py
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_verify.err
python -c "
import json
d=json.load(open('gpurun_out/bench_verify.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['kernel'])
for k,v in d['roofline']['kernels'].items(): print(' ',k,v)
for k in ('c1','c3','pairing','prepared','c4','c5'): print(k, json.dumps(d.get(k)))
print('cpu',d['cpu_baseline'])
"
