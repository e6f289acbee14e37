#!/bin/bash
# times every build/variants/lib*.so on the verify workload (and checks accept bits in warm-up)
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for so in build/variants/lib*.so; do
  v=$(basename $so .so)
  BN254_SO=$PWD/$so timeout 300 python bench.py --steps 3 --warmup 2 --sub none --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY' >> gpurun_out/variants.txt
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/var_%s.json'%v).read().strip().splitlines()[-1])
    r=d['roofline']
    print(v, "value=%.0f ms=%.2f"%(d['value'],d['ms_per_step']), {k.split('[')[0]:round(x['ms_per_launch'],2) for k,x in list(r['kernels'].items())[:4]})
except Exception as e:
    print(v,"FAILED",e, open('gpurun_out/var_%s.err'%v).read()[-300:])
PY
done
cat gpurun_out/variants.txt
