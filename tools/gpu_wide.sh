#!/bin/bash
# the wide slot machine: parity tests in every shape, then latency of small batches in the lane-pair and wide shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${K:-wide}" > gpurun_out/pytest_wide.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_wide.log)"; grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_wide.log | head -5
timeout 600 python - <<'PY' 2>&1 | tail -40
import os, sys, time
sys.path.insert(0, os.getcwd())
import bls_bn254_b200 as b
from bls_bn254_b200 import scheme
ctx = b.Context([0])
for n in (1, 256, 1024, 2048, 3072):
    pks, msgs, sigs, expect = scheme.synth_batch(ctx, n, variant=0, n_keys=min(n, 64), seed=3)
    row = []
    for lanes in (2, 4):
        ctx.set_option("lanes", lanes)
        ctx.verify_batch(0, pks, msgs, sigs, b.DST_MIN_SIG)
        ctx.profile(True); ctx.profile_drain()
        t0 = time.perf_counter()
        ok = ctx.verify_batch(0, pks, msgs, sigs, b.DST_MIN_SIG)
        dt = time.perf_counter() - t0
        recs = ctx.profile_drain(); ctx.profile(False)
        assert ok == expect, (n, lanes)
        ks = {k: round(ms, 3) for k, _, ms in recs if ms > 0.02}
        row.append("lanes=%d %.3f ms %s" % (lanes, dt * 1e3, ks))
    print("n=%d\n  %s\n  %s" % (n, row[0], row[1]))
PY
