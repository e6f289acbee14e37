#!/bin/bash
# hash-to-curve at 2^16 and 2^20 messages for every build/variants/lib*.so (CTA-size scan of the hash kernels)
mkdir -p gpurun_out
for so in build/variants/lib*.so; do
  for w in hash_g1 hash_g2; do for l in 16 20; do
    BN254_SO=$PWD/$so timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --log2n $l 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$(basename $so .so) $w 2^$l value %.0f ms %.3f' % (d['value'], d['ms_per_step']))"
  done; done
done
