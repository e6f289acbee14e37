#!/usr/bin/env python
"""BASELINE config C5: batch-verify 2^24 (pk, msg, sig) tuples through the host-buffer C ABI call
(bn254_verify_batch: chunks of 2^18 tuples staged H2D, verified, accept bytes copied back) on the
first --gpus devices of this box, one in-process context (one host thread per device).
The 2^24 tuples are a 2^--unique-tuple synthetic batch (1/64 corrupted) tiled; every accept byte is
checked against the construction.  Prints one JSON line per workload."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bls_bn254_b200 as b  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--unique", type=int, default=18, help="log2 of distinct tuples generated (tiled up to 2^log2n)")
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--group", type=int, default=256)
    args = ap.parse_args()
    n, u = 1 << args.log2n, 1 << min(args.unique, args.log2n)
    reps = n // u
    gen = b.Context([0])
    pks, msgs, sigs, exp = b.scheme.synth_batch(gen, u, variant=0, n_keys=4096)
    vpks, vmsgs, vsigs, _ = b.scheme.synth_batch(gen, u, variant=0, n_keys=4096, corrupt_every=0)
    gen.close()
    ctx = b.Context(list(range(args.gpus)))
    off = np.arange(n + 1, dtype=np.uint64) * 32
    seed = bytes(range(32))
    for name in ("verify", "verify_rand"):
        if name == "verify":
            P, M, S, want = pks * reps, b"".join(msgs) * reps, sigs * reps, exp * reps
            call = lambda: ctx.verify_batch(0, P, (M, off), S, b.DST_MIN_SIG)
        else:
            P, M, S = vpks * reps, b"".join(vmsgs) * reps, vsigs * reps
            want = b"\x01" * ((n + args.group - 1) // args.group)
            call = lambda: ctx.verify_batch_randomized(0, P, (M, off), S, b.DST_MIN_SIG, seed, group=args.group)
        ctx.verify_batch(0, P[:128 * 4096], (M[:32 * 4096], off[:4097]), S[:64 * 4096], b.DST_MIN_SIG)   # warm-up (module load, scratch)
        best = None
        for _ in range(args.passes):
            t0 = time.perf_counter()
            got = call()
            dt = time.perf_counter() - t0
            assert got == want, "accept bytes differ from the construction"
            best = dt if best is None else min(best, dt)
        print(json.dumps({"config": "C5: 2^%d tuples, min-sig, host buffers -> accept bytes, %d GPU(s), one process" % (args.log2n, args.gpus),
                          "workload": name, "n_gpus": args.gpus, "tuples": n, "seconds": best, "verifications_per_sec": n / best,
                          "h2d_bytes": len(P) + len(M) + len(S) + off.nbytes, "d2h_bytes": len(want),
                          "group": args.group if name == "verify_rand" else None, "accept_bytes_checked": True}), flush=True)
        del P, M, S
    ctx.close()


if __name__ == "__main__":
    main()
