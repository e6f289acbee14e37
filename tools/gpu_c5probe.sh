#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --log2n 18 --sub none --no-cpu-baseline > gpurun_out/b18.json 2> gpurun_out/b18.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/b18.json'))
print('2^18 value',d['value'],'e2e',d['e2e']['value'], d['clocks'])
for k,v in d['roofline']['kernels'].items(): print(' ',k,v)
"
for ch in 65536 262144 1048576; do
BN254_CHUNK=$ch python bench.py --steps 3 --warmup 3 --sub c5 --c5-log2n 22 --no-cpu-baseline > gpurun_out/c5_$ch.json 2> gpurun_out/c5_$ch.err; echo "rc=$?"
python -c "
import json,sys
d=json.load(open('gpurun_out/c5_$ch.json'))
print('chunk $ch', d['c5']['exact'], d['c5']['randomized'])
"
done
