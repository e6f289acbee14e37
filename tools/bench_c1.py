#!/usr/bin/env python
"""BASELINE config C1: sign + verify of ONE 32-byte message (one pairing check).  The reference crate would run this on the
CPU; timed here on the CPU port (oracle, one thread) and, for latency, through the GPU library with n = 1."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bls_bn254_b200 as b  # noqa: E402
from oracle import orc  # noqa: E402


def timeit(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    sk = b.scheme.scalar_from_seed(b.scheme.SEED, 7)
    msg = bytes(range(32))
    sk_b = sk.to_bytes(32, "big")
    # CPU port, one thread
    pk_c = orc.g2_mul(orc.g2_generator(), sk_b)

    def cpu_sign():
        return orc.g1_mul(orc.hash_to_g1([msg], b.DST_MIN_SIG), sk_b)

    sig_c = cpu_sign()

    def cpu_verify():
        assert orc.verify_batch(0, pk_c, [msg], sig_c, b.DST_MIN_SIG) == b"\x01"

    out = {"config": "C1: sign + verify of one 32-byte message, min-sig", "cpu_port_sign_ms": timeit(cpu_sign, 20), "cpu_port_verify_ms": timeit(cpu_verify, 20)}
    try:
        ctx = b.Context([0])
    except Exception as e:  # no GPU here
        out["gpu"] = "unavailable: %s" % e
        print(json.dumps(out))
        return
    pk_g = ctx.g2_mul(b.G2_GENERATOR, sk_b)

    def gpu_sign():
        return ctx.g1_mul(ctx.hash_to_g1([msg], b.DST_MIN_SIG), sk_b)

    sig_g = gpu_sign()
    assert (pk_g, sig_g) == (pk_c, sig_c), "GPU and CPU port disagree"

    def gpu_verify():
        assert ctx.verify_batch(0, pk_g, [msg], sig_g, b.DST_MIN_SIG) == b"\x01"

    out.update({"gpu_sign_ms": timeit(gpu_sign, 20), "gpu_verify_ms": timeit(gpu_verify, 20), "bytes_equal": True})
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
