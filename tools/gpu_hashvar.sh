#!/bin/bash
# hash-to-curve throughput of tuning builds (2^20 messages): BN254_SO selects the library
mkdir -p gpurun_out
for so in "$@"; do
  name=$(basename $so .so)
  if [ "$so" != "default" ]; then export BN254_SO=$PWD/$so; else unset BN254_SO; fi
  for w in hash_g1 hash_g2; do
  timeout 600 python bench.py --steps 3 --warmup 3 --sub none --workload $w --log2n 20 > gpurun_out/bench_${name}_$w.json 2> gpurun_out/bench_${name}_$w.err; rc=$?
  python -c "
import json
d=json.load(open('gpurun_out/bench_${name}_$w.json'))
print('$name $w rc=$rc value %.2f M/s e2e %.2f M/s; oracle-checked sample: %s'%(d['value']/1e6, d['e2e']['value']/1e6, d['cpu_baseline']['sample'][:40] if d['cpu_baseline'] else None))
"
  done
done
