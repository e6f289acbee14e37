#!/bin/bash
# one full ncu capture of the wide slot-machine kernels on a small batch (after a plain run of the same command)
mkdir -p gpurun_out
cat > /tmp/wide_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import bls_bn254_b200 as b
from bls_bn254_b200 import scheme
ctx = b.Context([0])
n = int(os.environ.get("WN", "16"))
pks, msgs, sigs, expect = scheme.synth_batch(ctx, n, variant=0, n_keys=min(n, 64), seed=3)
ctx.set_option("lanes", 4)
for _ in range(2):
    assert ctx.verify_batch(0, pks, msgs, sigs, b.DST_MIN_SIG) == expect
print("ok")
PY
CMD="python /tmp/wide_small.py"
$CMD > gpurun_out/plain_wide.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_w_" -s 2 -c 2 -o gpurun_out/${NCU_OUT:-prof_wide} -f $CMD > gpurun_out/ncu_wide.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_wide.log; cat gpurun_out/plain_wide.log | tail -2
