#!/bin/bash
# compare two builds across workloads/sizes
for so in build/variants/lib*.so; do
  for args in "--workload verify --log2n 16" "--workload verify --log2n 20 --steps 3" "--workload verify_pk --log2n 16" "--workload pairing --log2n 16" "--workload aggregate --log2n 16" "--workload verify --log2n 14" "--workload verify --log2n 18 --steps 3"; do
    BN254_SO=$PWD/$so python bench.py --warmup 2 --no-cpu-baseline --steps 5 $args 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $so)', '$args', 'value=%.0f ms=%.2f'%(d['value'],d['ms_per_step']))"
  done
done
