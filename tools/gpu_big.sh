#!/bin/bash
mkdir -p gpurun_out
for w in hash_g1 hash_g2 verify; do
  start=$(date +%s)
  python bench.py --steps 3 --warmup 3 --workload $w --log2n 20 --cpu-sample 4096 > gpurun_out/big_$w.json 2> gpurun_out/big_$w.err
  echo "$w rc=$? wall=$(( $(date +%s) - start ))s"; tail -2 gpurun_out/big_$w.err
  python - $w <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/big_%s.json'%w).read().strip().splitlines()[-1]); r=d["roofline"]
    print(d["config"]["tuples_per_gpu"], d["metric"], "value=%.0f ms=%.2f e2e=%.0f frac=%.3f whole=%.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["frac"],r["whole_step_frac"]))
except Exception as e: print("parse fail",e)
PY
done
