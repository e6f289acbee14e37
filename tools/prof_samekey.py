"""per-kernel times of the same-key randomised path (ctx.profile: a CUDA event pair around every launch)"""
import sys, time
import numpy as np
import bls_bn254_b200 as b

ctx = b.Context([0])
import os
SL = [int(x) for x in os.environ.get("SK_SLICES", "0").split(",")]
for log2n, nk, sl in [(16, 64, x) for x in SL] + [(18, 64, x) for x in SL] + [(16, 4096, 0)]:
    ctx.set_option("sk_msm_slices", sl)
    print("slices option", sl)
    n = 1 << log2n
    kp, kko, km, kof, ksg, kexp = b.scheme.synth_same_key_np(ctx, n, nk, 0, bad_keys=(1,))
    seed = bytes(range(32))
    cnt = [int(kko[i + 1] - kko[i]) for i in range(nk)]
    args = (0, kp.tobytes(), cnt, (km.tobytes(), kof), ksg.tobytes(), b.DST_MIN_SIG, seed)
    ctx.verify_same_key_randomized(*args)
    t0 = time.perf_counter()
    ok = ctx.verify_same_key_randomized(*args)
    dt = time.perf_counter() - t0
    assert ok == kexp.tobytes()
    ctx.profile(True); ctx.profile_drain()
    ctx.verify_same_key_randomized(*args)
    recs = ctx.profile_drain()
    ctx.profile(False)
    print("n=2^%d keys=%d: %.2f ms per call through the host API (%.2f M/s)" % (log2n, nk, dt * 1e3, n / dt / 1e6))
    agg = {}
    for r in recs:
        agg.setdefault(r[0], [0, 0.0])
        agg[r[0]][0] += 1
        agg[r[0]][1] += r[2]
    for k, v in agg.items():
        print("   %-28s x%-3d %8.3f ms" % (k, v[0], v[1]))
ctx.set_option("sk_msm_slices", 0)
# streaming: 2^22 tuples under 64 keys through the host call (chunks of 2^18 tuples, tails overlapped)
n = 1 << 22
kp, kko, km, kof, ksg, kexp = b.scheme.synth_same_key_np(ctx, n, 1024, 0, bad_keys=(1, 700))
cnt = [int(kko[i + 1] - kko[i]) for i in range(1024)]
args = (0, kp.tobytes(), cnt, (km.tobytes(), kof), ksg.tobytes(), b.DST_MIN_SIG, seed)
assert ctx.verify_same_key_randomized(*args) == kexp.tobytes()
t0 = time.perf_counter()
ok = ctx.verify_same_key_randomized(*args)
dt = time.perf_counter() - t0
assert ok == kexp.tobytes()
print("n=2^22 keys=1024 (64 keys per 2^18-tuple chunk): %.1f ms per call through the host API (%.2f M/s)" % (dt * 1e3, n / dt / 1e6))
ctx.close()
