#!/bin/bash
# verify workload at several batch sizes for every build/variants/lib*.so (CTA-size policy of the lane-pair kernels)
mkdir -p gpurun_out
for so in build/variants/lib*.so; do
  for l in ${SIZES:-13 14 15 16 17 18}; do
    BN254_SO=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --sub none --no-cpu-baseline --log2n $l 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$(basename $so .so) 2^$l value %.0f ms %.3f' % (d['value'], d['ms_per_step']), {k.split('[')[0]:round(v['ms_per_launch'],2) for k,v in list(d['roofline']['kernels'].items())[:3]})"
  done
done
