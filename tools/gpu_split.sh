#!/bin/bash
for sp in 1 2 3 4 6 8; do
  BN254_SPLIT=$sp python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('split', os.environ.get('SP'), 'value=%.0f ms=%.2f e2e=%.0f whole=%.3f kernel_pass=%.2f launches=%d'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['whole_step_frac'],r['kernel_pass_ms_per_step'],d['gpu_launches']))"
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
