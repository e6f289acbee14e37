#!/bin/bash
# randomised batch verification: parity tests, then the bench line at 2^16 / 2^20 and an ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "randomized" > gpurun_out/rbv_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/rbv_tests.log
for l in 16 20; do
  timeout 600 python bench.py --workload verify_rand --log2n $l --steps 5 --warmup 3 --cpu-sample 4096 > gpurun_out/rbv_bench_$l.json 2> gpurun_out/rbv_bench_$l.err
  echo "bench 2^$l rc=$?"; tail -2 gpurun_out/rbv_bench_$l.err
  python - $l <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/rbv_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value=%.0f ms=%.2f e2e=%.0f whole=%.3f cpu=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["whole_step_frac"],d["cpu_baseline"] and round(d["cpu_baseline"]["value"])))
    print({k:round(v,3) for k,v in r["kernels_ms_per_step"].items()})
except Exception as e: print("parse fail",e)
PY
done
