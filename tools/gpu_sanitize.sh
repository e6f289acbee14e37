#!/bin/bash
# compute-sanitizer on a small batch in the wide shape (racecheck: the wide machine orders its shared-memory traffic with
# __syncwarp) and memcheck on the fast-aggregate path.  TOOL=racecheck|memcheck
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import bls_bn254_b200 as b
from bls_bn254_b200 import scheme
ctx = b.Context([0])
pks, msgs, sigs, expect = scheme.synth_batch(ctx, 6, variant=0, n_keys=3, seed=3)
ctx.set_option("lanes", 4)
assert ctx.verify_batch(0, pks, msgs, sigs, b.DST_MIN_SIG) == expect
ctx.set_option("lanes", 0)
sks = [scheme.SecretKey(scheme.scalar_from_seed(5, i)) for i in range(3)]
keys = scheme.public_keys(ctx, sks, 0)
agg = scheme.aggregate(ctx, scheme.sign(ctx, sks, [b"m"] * 3, 0), 0)
assert ctx.fast_aggregate_verify_batch(0, keys, [3], [b"m"], agg, b.DST_MIN_SIG) == b"\x01"
print("ok")
PY
timeout 900 compute-sanitizer --tool ${TOOL:-racecheck} --print-limit 20 python /tmp/san_small.py > gpurun_out/sanitize_${TOOL:-racecheck}.log 2>&1
echo rc=$?; tail -15 gpurun_out/sanitize_${TOOL:-racecheck}.log
