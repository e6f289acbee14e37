#!/bin/bash
mkdir -p gpurun_out
for c in default 0 25 50 100; do
  if [ $c = default ]; then unset BN254_CARVEOUT; else export BN254_CARVEOUT=$c; fi
  for l in 16 20; do
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --log2n $l > gpurun_out/carve.json 2>gpurun_out/carve.err
    python - $c $l <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/carve.json').read().strip().splitlines()[-1]); r=d["roofline"]
    print("carveout", sys.argv[1], "2^"+sys.argv[2], "value=%.0f ms=%.2f"%(d["value"],d["ms_per_step"]), {k:round(v,2) for k,v in list(r["kernels_ms_per_step"].items())[:3]})
except Exception as e: print("fail", e, open('gpurun_out/carve.err').read()[-300:])
PY
  done
done
