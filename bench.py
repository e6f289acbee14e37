#!/usr/bin/env python3
"""bench.py -- BLS-BN254 batch verification throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload verify|verify_pk|verify_rand|hash_g1|hash_g2|pairing|aggregate]
                  [--sub all|none|c1,c3,c4,c5,pairing]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...       # the CPU arm: the oracle port on all host cores

A step = one pass of the hot path (hash-to-G1 -> 2-pair Miller loop -> final exponentiation -> accept bit) over one
batch of 2^16 synthetic (pk, msg, sig) tuples PER GPU (BASELINE config 2; weak scaling: every rank owns a disjoint shard,
no inter-GPU traffic).  `value` is timed with CUDA events on the library's launch stream with all inputs resident in
HBM; `e2e` is the same batch through the host-buffer C ABI call (pinned host buffers, H2D + kernels + D2H inside the
timed region).  The same JSON line carries one sub-record per remaining BASELINE config, measured in the same run:
  c1       sign + verify of ONE 32-byte message (latency; GPU library and the CPU port)
  c3       hash_to_curve: 2^20 messages per GPU to G1 and to G2
  pairing  2^16 standalone pairings per GPU (bn254_pairing_batch: Miller loop + final exponentiation) -> pairings_per_sec
  c4       aggregate verify: 2^16 signatures in TOTAL sharded over the N ranks, ONE ncclAllGather inside the library,
           one final exponentiation (strong scaling)
  c5       batch verify 2^24 tuples in TOTAL sharded over the N ranks through the host-buffer call (strong scaling),
           exact and randomised
  prepared the primary batch again with its 4096 public keys PREPARED once (G2Prepared, bn254_verify_batch_prepared)
Rank 0 prints ONE JSON line.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# canonical work per unit in Fp multiplications M; 1 M = 264 INT32 IMAD-ops
M_IMAD = 264
# verify_rand (randomised batch verification, groups of 256, 64-bit scalars a + b lambda): hash 2204 + one-pair Miller loop 8610 +
# two joint double-and-add scalings (32 x 8 + 25 x 12 + 1 = 557 each) + one affine conversion 332 + product tree 58 + per-group
# pairing / 256 = 68
WORK_M = {"verify": 23015, "verify_pk": 29687, "hash_g1": 2204, "hash_g2": 6400, "pairing": 17037, "aggregate": 10868, "verify_rand": 12386}
KERNEL_M = {"k_miller2_fixed": 12384, "k_miller2": 14860, "k_miller_twin": 14860, "k_miller1": 8610, "k_final_exp": 8427, "k_hash_g1": 2204, "k_hash_g2": 6400,
            "k_rbv_scale_g1": 889, "k_rbv_scale_g1p": 557}
UNIT = {"verify": "verifications/s", "verify_pk": "verifications/s", "hash_g1": "hashes/s", "hash_g2": "hashes/s", "pairing": "pairings/s",
        "aggregate": "verifications/s", "verify_rand": "verifications/s"}
METRIC = {"verify": "bls_verifications_per_sec", "verify_pk": "bls_verifications_per_sec", "hash_g1": "hash_to_g1_per_sec",
          "hash_g2": "hash_to_g2_per_sec", "pairing": "pairings_per_sec", "aggregate": "bls_aggregate_verified_signatures_per_sec",
          "verify_rand": "bls_randomized_batch_verifications_per_sec"}
RBV_SEED = hashlib.sha256(b"bench.py randomised batch verification seed").digest()
N_KEYS = 4096          # distinct keys of every synthetic batch
CORRUPT_EVERY = 64     # one tuple in 64 is broken, so both verdicts occur


def executed_ops():
    """EXECUTED multiplier work per item (INT32 multiply-add ops), frozen by tools/count_ops.py from the host build"""
    try:
        return {k: v["imad_ops_per_item"] for k, v in json.load(open(os.path.join(ROOT, "profiles", "executed_ops.json")))["kernels"].items()}
    except Exception:
        return {}


def source_sha16():
    """identity of the kernel sources: ncu-derived numbers committed under profiles/ are only quoted for the sources they were taken from"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "bls_bn254_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent sampling (NVML) of SM clock and throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def make_messages(n, first):
    from bls_bn254_b200 import scheme
    m = scheme.synth_messages_np(n, first)
    return [bytes(r) for r in m]


class OracleGroupOps:
    """The group operations scheme.synth_batch needs (keys, signatures), on the CPU oracle: the reference arm builds the SAME
    synthetic batch as the GPU arm without touching the CUDA library."""

    def __init__(self, orc, threads):
        self.o, self.t = orc, threads

    def g1_mul(self, pts, sc): return self.o.g1_mul(pts, sc, nthreads=self.t)
    def g2_mul(self, pts, sc): return self.o.g2_mul(pts, sc, nthreads=self.t)
    def hash_to_g1(self, msgs, dst): return self.o.hash_to_g1(msgs, dst, nthreads=self.t)
    def hash_to_g2(self, msgs, dst): return self.o.hash_to_g2(msgs, dst, nthreads=self.t)


def run_reference(args):
    """CPU arm: the oracle port of the reference algorithms on all host cores (the Rust crate cannot be built in this image:
    no cargo/rustc).  Each step runs the FIRST --ref-sample tuples of the very batch the GPU arm verifies (same seed, same
    4096 keys, same corrupted tuples), and the accept bits are checked against the construction."""
    import bls_bn254_b200 as b
    from oracle import orc
    cores = orc.ncpu()
    n = min(args.ref_sample, 1 << args.log2n)
    w = args.workload
    ops = OracleGroupOps(orc, cores)
    variant = b.MIN_PK if w == "verify_pk" else b.MIN_SIG
    dst = b.DST_MIN_PK if w in ("verify_pk", "hash_g2") else b.DST_MIN_SIG
    rpk, rsig = (128, 64) if variant == 0 else (64, 128)
    expected = pks = sigs = agg = g1 = g2 = None
    if w in ("verify", "verify_pk"):
        pks, msgs, sigs, expected = b.scheme.synth_batch(ops, n, variant=variant, n_keys=N_KEYS, corrupt_every=CORRUPT_EVERY)
    elif w in ("verify_rand", "aggregate"):
        pks, msgs, sigs, _ = b.scheme.synth_batch(ops, n, variant=variant, n_keys=N_KEYS, corrupt_every=0)
        agg = orc.g1_sum(sigs)
    else:
        msgs = make_messages(n, 0)
        if w == "pairing":
            g1, g2 = ops.hash_to_g1(msgs, b.DST_MIN_SIG), ops.hash_to_g2(msgs, b.DST_MIN_PK)

    def step():
        if w in ("verify", "verify_pk"):
            assert orc.verify_batch(variant, pks, msgs, sigs, dst, nthreads=cores) == expected, "oracle and construction disagree"
        elif w == "verify_rand":
            assert all(orc.verify_batch_randomized(0, pks, msgs, sigs, dst, RBV_SEED, group=args.group, nthreads=cores))
        elif w == "hash_g1":
            orc.hash_to_g1(msgs, dst, nthreads=cores)
        elif w == "hash_g2":
            orc.hash_to_g2(msgs, dst, nthreads=cores)
        elif w == "pairing":
            orc.pairing_batch(g1, g2, nthreads=cores)
        else:
            assert orc.aggregate_verify(0, pks, msgs, agg, dst, nthreads=cores) == 1

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = ("the first %d tuples of the arm's own synthetic batch per step (same seed, %d keys, %s), oracle C port of the reference algorithms "
              "(4x64-bit Montgomery), %d threads; accept bits checked against the construction" %
              (n, N_KEYS, "1 tuple in %d corrupted" % CORRUPT_EVERY if expected is not None else "all tuples valid", cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC[w], "value": v, "unit": UNIT[w], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(w, args),
        "cpu_baseline": {"value": v, "unit": UNIT[w], "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT[w], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(w, args):
    names = {
        "verify": "batch verify 2^%d independent (pk, msg, sig) tuples per GPU, min-sig (pk in G2, sig in G1), distinct 32-byte messages",
        "verify_pk": "batch verify 2^%d independent tuples per GPU, min-pk (pk in G1, sig in G2), distinct 32-byte messages",
        "hash_g1": "hash_to_curve: 2^%d 32-byte messages per GPU to G1 (expand_message_xmd SHA-256 + SvdW)",
        "hash_g2": "hash_to_curve: 2^%d 32-byte messages per GPU to G2 (expand_message_xmd SHA-256 + SvdW + cofactor clearing)",
        "pairing": "2^%d pairings per GPU (Miller loop + final exponentiation), distinct random G1/G2 points",
        "aggregate": "aggregate verify: 2^%d signatures on distinct messages per GPU, one multi-Miller product + single final exponentiation",
        "verify_rand": "randomised batch verify of 2^%d valid (pk, msg, sig) tuples per GPU, min-sig, one accept byte per group of "
                       + str(args.group) + " tuples (64-bit scalars, one final exponentiation per group)",
    }
    return {"workload": names[w] % args.log2n, "tuples_per_gpu": 1 << args.log2n, "variant": "min-pk" if w == "verify_pk" else "min-sig",
            "msg_bytes": 32, "distinct_keys": N_KEYS,
            "input_validation": "on-curve + subgroup checks inside the timed region" if getattr(args, "check_subgroup", False) else "none (trusted inputs); identity keys / signatures rejected",
            "corrupted_fraction": "1/%d" % CORRUPT_EVERY if w in ("verify", "verify_pk") else "0",
            "l2": "L2 flushed between timed steps (256 MiB write)", "sharding": "independent tuples by index, no data-path collective"
            if w != "aggregate" else "pairs by index; one ncclAllGather of 400-byte records inside the library"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="verify", choices=sorted(WORK_M))
    ap.add_argument("--log2n", type=int, default=16, help="log2 of tuples per GPU per step")
    ap.add_argument("--group", type=int, default=256, help="tuples per accept byte of the verify_rand workload")
    ap.add_argument("--ref-sample", type=int, default=16384, help="tuples per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=65536, help="tuples of the cpu_baseline leg (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--check-subgroup", action="store_true", help="verify workloads: also validate inputs (on-curve + order-r subgroup)")
    ap.add_argument("--same-key-keys", type=int, default=64, help="distinct keys of the same_key sub-record")
    ap.add_argument("--same-key-stream-log2n", type=int, default=22, help="log2 of the tuples per GPU of the streamed same_key figure")
    ap.add_argument("--sub", default=None, help="sub-records of the other BASELINE configs: all | none | comma list of c1,c3,c4,c5,pairing,prepared "
                                                "(default: all for the verify workload, none otherwise)")
    ap.add_argument("--c5-log2n", type=int, default=24, help="log2 of the TOTAL tuples of the c5 sub-record (sharded over the ranks)")
    ap.add_argument("--c3-log2n", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import bls_bn254_b200 as b

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bls_bn254_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = args.workload
    subs = args.sub if args.sub is not None else ("all" if w == "verify" else "none")
    subs = {"all": {"c1", "c3", "c4", "c5", "pairing", "prepared", "same_key"}, "none": set()}.get(subs, set(subs.split(",")))
    variant = b.MIN_PK if w == "verify_pk" else b.MIN_SIG
    vflags = b.CHECK_SUBGROUP if args.check_subgroup else 0
    n = 1 << args.log2n
    ctx = b.Context([local_rank])
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    if world > 1:
        b.dist.init_comm(ctx, device=dev)   # the library's own NCCL communicator (aggregate path)

    def dev_bytes(x):
        return torch.frombuffer(bytearray(x) if isinstance(x, (bytes, bytearray)) else x, dtype=torch.uint8).to(dev)

    def pinned(x):
        return torch.frombuffer(bytearray(x) if isinstance(x, (bytes, bytearray)) else x, dtype=torch.uint8).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_dev(step, steps):
        """steps passes of `step` (everything enqueued on the library stream), CUDA events on that stream, L2 flushed between
        passes; returns (sum of device ms, sum of host wall ms)"""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall = 0.0
        for k in range(steps):
            flush.fill_(k & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            evs[k][0].record(stream)
            step()
            evs[k][1].record(stream)
            evs[k][1].synchronize()
            wall += (time.perf_counter() - t0) * 1e3
        return sum(a.elapsed_time(bb) for a, bb in evs), wall

    # ---- synthetic inputs of the primary workload (generated on the GPU, untimed) --------------------------
    first = rank * n
    dst = b.DST_MIN_PK if w in ("verify_pk", "hash_g2") else b.DST_MIN_SIG
    pks = sigs = agg = g1 = g2 = None
    expected = None
    if w in ("verify", "verify_pk"):
        pks, msgs, sigs, expected = b.scheme.synth_batch(ctx, n, variant=variant, first=first, n_keys=min(n, N_KEYS), corrupt_every=CORRUPT_EVERY)
    elif w == "verify_rand":
        pks, msgs, sigs, _ = b.scheme.synth_batch(ctx, n, variant=variant, first=first, n_keys=min(n, N_KEYS), corrupt_every=0)
        ngroups = (n + args.group - 1) // args.group
        expected = b"\x01" * ngroups
    elif w == "aggregate":
        pks, msgs, sigs, _ = b.scheme.synth_batch(ctx, n, variant=variant, first=first, n_keys=min(n, N_KEYS), corrupt_every=0)
        agg_local = b.scheme.aggregate(ctx, sigs, variant)
    else:
        msgs = make_messages(n, first)
        if w == "pairing":
            g1 = ctx.hash_to_g1(msgs, b.DST_MIN_SIG)
            g2 = ctx.hash_to_g2(msgs, b.DST_MIN_PK)
    blob, off = b.pack_msgs(msgs)

    d_blob = dev_bytes(blob if blob else b"\0")
    d_off = torch.tensor(list(off), dtype=torch.int64, device=dev)
    d_dst = dev_bytes(dst)
    d_pks = dev_bytes(pks) if pks else None
    d_sigs = dev_bytes(sigs) if sigs else None
    d_g1 = dev_bytes(g1) if g1 else None
    d_g2 = dev_bytes(g2) if g2 else None
    out_rec = {"verify": 1, "verify_pk": 1, "hash_g1": 64, "hash_g2": 128, "pairing": 384, "aggregate": 1, "verify_rand": 1}[w]
    d_out = torch.zeros({"aggregate": 16, "verify_rand": (n + args.group - 1) // args.group}.get(w, n * out_rec), dtype=torch.uint8, device=dev)
    d_seed = dev_bytes(RBV_SEED)
    d_status = torch.zeros(n, dtype=torch.uint8, device=dev)

    def job_aggregate_signature(local_sum):
        """sum over the ranks of the per-rank signature sums (setup, untimed): what rank 0 hands to the aggregate check"""
        if world == 1:
            return local_sum
        t = dev_bytes(local_sum)
        allp = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allp, t)
        return ctx.g1_sum(b"".join(bytes(x.cpu().numpy()) for x in allp))

    if w == "aggregate":
        agg = job_aggregate_signature(agg_local)
        d_agg = dev_bytes(agg) if rank == 0 else None

    def step_dev():
        """device-resident pass: everything is enqueued on the library stream"""
        if w in ("verify", "verify_pk"):
            ctx.verify_batch_dev(variant, d_pks.data_ptr(), d_blob.data_ptr(), d_off.data_ptr(), d_sigs.data_ptr(), n, d_dst.data_ptr(),
                                 len(dst), vflags, d_out.data_ptr())
        elif w == "verify_rand":
            ctx.verify_batch_randomized_dev(variant, d_pks.data_ptr(), d_blob.data_ptr(), d_off.data_ptr(), d_sigs.data_ptr(), n, d_dst.data_ptr(),
                                            len(dst), vflags, d_seed.data_ptr(), 64, args.group, d_out.data_ptr())
        elif w == "hash_g1":
            ctx.hash_to_g1_dev(d_blob.data_ptr(), d_off.data_ptr(), n, d_dst.data_ptr(), len(dst), 2, d_out.data_ptr(), d_status.data_ptr())
        elif w == "hash_g2":
            ctx.hash_to_g2_dev(d_blob.data_ptr(), d_off.data_ptr(), n, d_dst.data_ptr(), len(dst), 2, d_out.data_ptr(), d_status.data_ptr())
        elif w == "pairing":
            ctx.pairing_dev(d_g1.data_ptr(), d_g2.data_ptr(), n, d_out.data_ptr(), d_status.data_ptr())
        else:
            # shards -> product tree -> ONE ncclAllGather (400 B per rank) -> ONE post-gather kernel, all on the library stream
            ctx.aggregate_verify_dist_dev(variant, d_pks.data_ptr(), d_blob.data_ptr(), d_off.data_ptr(), n, d_agg.data_ptr() if rank == 0 else None,
                                          d_dst.data_ptr(), len(dst), vflags, d_out.data_ptr())

    # e2e inputs live in PINNED host memory and go through the host-buffer C entry points (Context.raw): every step copies
    # them to the device, runs the kernels and copies the result back inside the timed region
    h_blob = pinned(blob if blob else b"\0")
    h_off = torch.tensor(list(off), dtype=torch.int64).pin_memory()
    h_pks = pinned(pks) if pks else None
    h_sigs = pinned(sigs) if sigs else None
    h_ok = torch.zeros(max(n, 1), dtype=torch.uint8).pin_memory()
    h_g1 = pinned(g1) if g1 else None
    h_g2 = pinned(g2) if g2 else None
    h_out = torch.zeros(n * out_rec, dtype=torch.uint8).pin_memory() if w in ("hash_g1", "hash_g2", "pairing") else None

    def step_e2e():
        """the call a user makes: host buffers in, host result out"""
        if w in ("verify", "verify_pk"):
            ctx.raw("bn254_verify_batch", variant, h_pks.data_ptr(), h_blob.data_ptr(), h_off.data_ptr(), h_sigs.data_ptr(), n, dst, len(dst), vflags,
                    h_ok.data_ptr())
            return bytes(h_ok[:n].numpy())
        if w == "verify_rand":
            ctx.raw("bn254_verify_batch_randomized", variant, h_pks.data_ptr(), h_blob.data_ptr(), h_off.data_ptr(), h_sigs.data_ptr(), n, dst,
                    len(dst), vflags, RBV_SEED, 64, args.group, h_ok.data_ptr())
            return bytes(h_ok[:(n + args.group - 1) // args.group].numpy())
        if w in ("hash_g1", "hash_g2"):
            ctx.raw("bn254_hash_to_g1" if w == "hash_g1" else "bn254_hash_to_g2", h_blob.data_ptr(), h_off.data_ptr(), n, dst, len(dst),
                    h_out.data_ptr(), h_ok.data_ptr())
            return None
        if w == "pairing":
            ctx.raw("bn254_pairing_batch", h_g1.data_ptr(), h_g2.data_ptr(), n, h_out.data_ptr(), h_ok.data_ptr())
            return None
        return ctx.aggregate_verify_dist(variant, pks, (blob, off), agg if rank == 0 else None, dst, vflags)

    # ---- roofline denominator: INT32 multiply-add peak, measured live --------------------------
    peaks, field_peaks = {}, {}
    if rank == 0:
        for mode, name in ((0, "mad.lo"), (1, "mad.lo+mad.hi"), (2, "imad.wide")):
            peaks[name] = max(ctx.imad_peak(mode, 20000)[0] for _ in range(2))
        # field-level ceilings (registers only, no memory): canonical IMAD-ops/s of dependent Montgomery products /
        # fused dot products with 1 and 4 warps per scheduler
        for mode, name in ((1, "fp_mul"), (0, "fp_dot2")):
            for wps in (4, 16):
                field_peaks["%s_%dwarps_per_sm" % (name, wps)] = ctx.fp_peak(mode, wps, 4000) / 1e12

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        step_dev()
        ctx.sync()
    if args.warmup:
        got = step_e2e()
        if expected is not None:
            assert got == expected, "accept bits differ from the construction"
            assert bytes(d_out.cpu().numpy()) == expected, "device-resident accept bits differ from the construction"
        if w == "aggregate":
            assert got == 1 and int(d_out[0].item()) == 1, "aggregate verification rejected a valid aggregate"
        if h_out is not None:
            assert bytes(h_out.numpy()) == bytes(d_out.cpu().numpy()), "host-buffer call and device-resident call disagree"

    # ---- timed region: EXACTLY K steps, CUDA events on the launch stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    barrier()
    sampler.start()
    dev_ms, host_ms = timed_dev(step_dev, args.steps)
    barrier()
    clocks = sampler.result()
    launches = ctx.launch_count - launches0
    total_ms = max_over_ranks(dev_ms)
    value = n * world * args.steps / (total_ms * 1e-3)

    # ---- per-kernel durations for the roofline: the same K steps again with a CUDA event pair around
    #      every launch (the library then runs its kernels back to back on one stream)
    ctx.profile(True)
    ctx.profile_drain()
    prof_ms, _ = timed_dev(step_dev, args.steps)
    prof = ctx.profile_drain()
    ctx.profile(False)
    barrier()

    # ---- e2e: host buffers through the public call, copies inside the timed region --------------
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    in_bytes = {"verify": len(blob) + 8 * (n + 1) + n * 192 + len(dst), "verify_pk": len(blob) + 8 * (n + 1) + n * 192 + len(dst),
                "hash_g1": len(blob) + 8 * (n + 1) + len(dst), "hash_g2": len(blob) + 8 * (n + 1) + len(dst), "pairing": n * 192,
                "aggregate": len(blob) + 8 * (n + 1) + n * 128 + 64 + len(dst), "verify_rand": len(blob) + 8 * (n + 1) + n * 192 + len(dst) + 32}[w]
    out_bytes = {"verify": n, "verify_pk": n, "hash_g1": n * 65, "hash_g2": n * 129, "pairing": n * 385, "aggregate": 1,
                 "verify_rand": (n + args.group - 1) // args.group}[w]

    # =====================================================================================================
    # sub-records: the other BASELINE configs, same run, every rank takes part
    # =====================================================================================================
    sub = {}
    ssteps = max(2, min(args.steps, 3))

    if "prepared" in subs and w == "verify":
        # the primary batch once more, its N_KEYS distinct public keys prepared ONCE (G2Prepared): tuple i names key i mod N_KEYS
        nk = min(n, N_KEYS)
        upk = b.scheme.public_keys(ctx, [b.scheme.SecretKey(b.scheme.scalar_from_seed(b.scheme.SEED, i)) for i in range(nk)], 0)   # synth_batch's key list
        t0 = time.perf_counter()
        prep = ctx.g2_prepare(upk)
        prep_s = time.perf_counter() - t0
        kidx = np.arange(n, dtype=np.uint32) % np.uint32(nk)
        # synth_batch replaces the key of some corrupted tuples by a neighbour's: follow that in the index array
        for j in range(n):
            if expected[j] == 0 and pks[128 * j:128 * (j + 1)] != upk[128 * (j % nk):128 * (j % nk + 1)]:
                kidx[j] = (j + 1) % nk
        dki = torch.from_numpy(kidx.view(np.int32)).to(dev)
        hki = torch.from_numpy(kidx.view(np.int32)).pin_memory()
        dokp = torch.zeros(n, dtype=torch.uint8, device=dev)
        run = lambda: ctx.verify_batch_prepared_dev(prep, dki.data_ptr(), d_blob.data_ptr(), d_off.data_ptr(), d_sigs.data_ptr(), n, d_dst.data_ptr(), len(dst),
                                                    vflags, dokp.data_ptr())
        run(); ctx.sync()
        assert bytes(dokp.cpu().numpy()) == expected, "prepared-key verification differs from the construction"
        barrier()
        ms, _ = timed_dev(run, args.steps)
        ms = max_over_ranks(ms)
        ctx.raw("bn254_verify_batch_prepared", prep._h, hki.data_ptr(), h_blob.data_ptr(), h_off.data_ptr(), h_sigs.data_ptr(), n, dst, len(dst), vflags, h_ok.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.raw("bn254_verify_batch_prepared", prep._h, hki.data_ptr(), h_blob.data_ptr(), h_off.data_ptr(), h_sigs.data_ptr(), n, dst, len(dst), vflags, h_ok.data_ptr())
        ems = max_over_ranks((time.perf_counter() - t0) * 1e3)
        assert bytes(h_ok[:n].numpy()) == expected
        sub["prepared"] = {"config": "the primary batch with its %d distinct public keys prepared once (G2Prepared: 88 line triples per key, %.1f MB on the "
                                     "device); bn254_verify_batch_prepared: both Miller line sequences from tables, no twist-point arithmetic" % (nk, nk * 16896 / 1e6),
                           "verifications_per_sec": n * world * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                           "e2e_verifications_per_sec": n * world * args.steps / (ems * 1e-3), "prepare_seconds": prep_s,
                           "checked": "accept bytes equal the construction (= the exact path's)"}
        prep.close()
        del dki, dokp

    if "pairing" in subs:
        pm = make_messages(n, first + (1 << 40))
        pg1, pg2 = ctx.hash_to_g1(pm, b.DST_MIN_SIG), ctx.hash_to_g2(pm, b.DST_MIN_PK)
        dp1, dp2 = dev_bytes(pg1), dev_bytes(pg2)
        dgt = torch.zeros(n * 384, dtype=torch.uint8, device=dev)
        hp1, hp2 = pinned(pg1), pinned(pg2)
        hgt = torch.zeros(n * 384, dtype=torch.uint8).pin_memory()
        hst = torch.zeros(n, dtype=torch.uint8).pin_memory()
        run = lambda: ctx.pairing_dev(dp1.data_ptr(), dp2.data_ptr(), n, dgt.data_ptr(), d_status.data_ptr())
        run(); ctx.sync()
        barrier()
        ms, _ = timed_dev(run, ssteps)
        ms = max_over_ranks(ms)
        ctx.raw("bn254_pairing_batch", hp1.data_ptr(), hp2.data_ptr(), n, hgt.data_ptr(), hst.data_ptr())   # untimed: sizes the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(ssteps):
            ctx.raw("bn254_pairing_batch", hp1.data_ptr(), hp2.data_ptr(), n, hgt.data_ptr(), hst.data_ptr())
        ems = max_over_ranks((time.perf_counter() - t0) * 1e3)
        assert bytes(hgt.numpy()) == bytes(dgt.cpu().numpy()), "pairing: host-buffer call and device-resident call disagree"
        check = None
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import orc
            assert orc.pairing_batch(pg1[:64 * 64], pg2[:128 * 64], nthreads=orc.ncpu()) == bytes(hgt[:384 * 64].numpy()), "pairing differs from the oracle"
            check = "first 64 GT elements byte-equal to the oracle"
        sub["pairing"] = {"config": "2^%d standalone pairings per GPU (bn254_pairing_batch: one Miller loop + one final exponentiation each)" % args.log2n,
                          "pairings_per_sec": n * world * ssteps / (ms * 1e-3), "ms_per_step": ms / ssteps, "steps": ssteps,
                          "e2e_pairings_per_sec": n * world * ssteps / (ems * 1e-3), "checked": check}
        del dp1, dp2, dgt
        torch.cuda.empty_cache()

    if "same_key" in subs:
        # SURVEY 8(f)-1 "same-key grouping": the batch size of the primary workload, signed under FEW keys (tuples sorted by key);
        # bn254_verify_same_key_randomized gives one verdict per key from one Miller loop + one final exponentiation per KEY
        nks = args.same_key_keys
        bad_keys = (1, nks // 2) if nks > 2 else ()
        kp, kko, km, kof, ksg, kexp = b.scheme.synth_same_key_np(ctx, n, nks, 0, first=first + (1 << 43), bad_keys=bad_keys)
        kseed = bytes((7 * i + 1) & 255 for i in range(32))
        dkp, dkm, dks = dev_bytes(kp.reshape(-1)), dev_bytes(km.reshape(-1)), dev_bytes(ksg.reshape(-1))
        dkof = torch.from_numpy(kof.view(np.int64)).to(dev)
        dkseed, dkdst = dev_bytes(kseed), dev_bytes(b.DST_MIN_SIG)
        dkok = torch.zeros(nks, dtype=torch.uint8, device=dev)
        hkp, hkm, hks = pinned(kp.reshape(-1)), pinned(km.reshape(-1)), pinned(ksg.reshape(-1))
        hkof = torch.from_numpy(kof.view(np.int64)).pin_memory()
        hkok = torch.zeros(nks, dtype=torch.uint8).pin_memory()
        run = lambda: ctx.verify_same_key_randomized_dev(0, dkp.data_ptr(), kko.ctypes.data, nks, dkm.data_ptr(), dkof.data_ptr(), dks.data_ptr(),
                                                         dkdst.data_ptr(), len(b.DST_MIN_SIG), vflags, dkseed.data_ptr(), 64, dkok.data_ptr())
        run(); ctx.sync()
        assert bytes(dkok.cpu().numpy()) == kexp.tobytes(), "same-key verdicts differ from the construction"
        barrier()
        ms, _ = timed_dev(run, args.steps)
        ms = max_over_ranks(ms)
        erun = lambda: ctx.raw("bn254_verify_same_key_randomized", 0, hkp.data_ptr(), kko.ctypes.data, nks, hkm.data_ptr(), hkof.data_ptr(), hks.data_ptr(),
                               b.DST_MIN_SIG, len(b.DST_MIN_SIG), vflags, kseed, 64, hkok.data_ptr())
        erun()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            erun()
        ems = max_over_ranks((time.perf_counter() - t0) * 1e3)
        assert bytes(hkok.numpy()) == kexp.tobytes(), "same-key verdicts (host-buffer call) differ from the construction"
        check = None
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import orc
            kk = min(nks, 4)                       # the first keys (a broken one among them) against the oracle
            tk = int(kko[kk])
            cnt = [int(kko[i + 1] - kko[i]) for i in range(kk)]
            sub_ok = ctx.verify_same_key_randomized(0, kp[:kk].tobytes(), cnt, (km[:tk].tobytes(), kof[:tk + 1]), ksg[:tk].tobytes(), b.DST_MIN_SIG, kseed, flags=vflags)
            ref = orc.verify_same_key_randomized(0, kp[:kk].tobytes(), cnt, [bytes(r) for r in km[:tk]], ksg[:tk].tobytes(), b.DST_MIN_SIG, kseed, flags=vflags,
                                                 nthreads=orc.ncpu())
            assert sub_ok == ref == kexp[:kk].tobytes(), "same-key verdicts differ from the oracle"
            check = "verdicts equal the construction; the first %d keys (%d tuples, one broken key) equal the oracle" % (kk, tk)
        sub["same_key"] = {"config": "2^%d (msg, sig) tuples per GPU under %d keys, sorted by key (min-sig, %d keys with one broken tuple); "
                                     "bn254_verify_same_key_randomized: per tuple a hash to G1 and a share of the key's two sums (bucket-method "
                                     "multi-scalar multiplications over 64-bit GLV scalars), per KEY one two-pair Miller loop and one final exponentiation; one accept byte per key" % (args.log2n, nks, len(bad_keys)),
                           "verifications_per_sec": n * world * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                           "e2e_verifications_per_sec": n * world * args.steps / (ems * 1e-3), "keys": nks, "rand_bits": 64, "checked": check}
        del dkp, dkm, dks, dkof, dkok
        torch.cuda.empty_cache()
        # streamed: 2^22 tuples per GPU under 16 x as many keys through the host call on pinned buffers -- chunks of 2^18 tuples, the
        # per-key tail of a chunk (a latency-bound chain on a few keys) overlaps the per-tuple kernels of the next
        n_st, nk_st = (1 << args.same_key_stream_log2n) if args.same_key_stream_log2n else 0, 16 * nks
        if n_st:
            sp, sko, sm, sof, ssg, sexp = b.scheme.synth_same_key_np(ctx, n_st, nk_st, 0, first=first + (1 << 44), bad_keys=(3, nk_st - 2))
            hsp, hsm, hss = pinned(sp.reshape(-1)), pinned(sm.reshape(-1)), pinned(ssg.reshape(-1))
            hsof = torch.from_numpy(sof.view(np.int64)).pin_memory()
            hsok = torch.zeros(nk_st, dtype=torch.uint8).pin_memory()
            srun = lambda: ctx.raw("bn254_verify_same_key_randomized", 0, hsp.data_ptr(), sko.ctypes.data, nk_st, hsm.data_ptr(), hsof.data_ptr(), hss.data_ptr(),
                                   b.DST_MIN_SIG, len(b.DST_MIN_SIG), vflags, kseed, 64, hsok.data_ptr())
            srun()
            assert bytes(hsok.numpy()) == sexp.tobytes(), "same-key verdicts (streamed) differ from the construction"
            barrier()
            t0 = time.perf_counter()
            srun()
            torch.cuda.synchronize()
            st_s = max_over_ranks(time.perf_counter() - t0)
            sub["same_key"]["streamed"] = {"tuples_per_gpu": n_st, "keys_per_gpu": nk_st, "seconds": st_s, "verifications_per_sec": n_st * world / st_s,
                                           "h2d_bytes_per_gpu": int(hsp.numel() + hsm.numel() + hss.numel() + 8 * (n_st + 1)), "d2h_bytes_per_gpu": nk_st,
                                           "note": "host-buffer call, copies inside the timed region; verdicts equal the construction"}
            del hsp, hsm, hss, hsof

    if "c3" in subs:
        n3 = 1 << args.c3_log2n
        m3 = b.scheme.synth_messages_np(n3, rank * n3 + (1 << 41))
        off3 = np.arange(0, 32 * (n3 + 1), 32, dtype=np.uint64)
        dm3, do3 = torch.from_numpy(m3.reshape(-1)).to(dev), torch.from_numpy(off3.view(np.int64)).to(dev)
        hm3, ho3 = torch.from_numpy(m3.reshape(-1)).pin_memory(), torch.from_numpy(off3.view(np.int64)).pin_memory()
        rec3 = {}
        for grp, rec, dstg in ((1, 64, b.DST_MIN_SIG), (2, 128, b.DST_MIN_PK)):
            dd = dev_bytes(dstg)
            dout = torch.zeros(n3 * rec, dtype=torch.uint8, device=dev)
            dst3 = torch.zeros(n3, dtype=torch.uint8, device=dev)
            hout = torch.zeros(n3 * rec, dtype=torch.uint8).pin_memory()
            hst3 = torch.zeros(n3, dtype=torch.uint8).pin_memory()
            fn = ctx.hash_to_g1_dev if grp == 1 else ctx.hash_to_g2_dev
            run = lambda: fn(dm3.data_ptr(), do3.data_ptr(), n3, dd.data_ptr(), len(dstg), 2, dout.data_ptr(), dst3.data_ptr())
            run(); ctx.sync()
            barrier()
            ms, _ = timed_dev(run, ssteps)
            ms = max_over_ranks(ms)
            ctx.raw("bn254_hash_to_g1" if grp == 1 else "bn254_hash_to_g2", hm3.data_ptr(), ho3.data_ptr(), n3, dstg, len(dstg), hout.data_ptr(), hst3.data_ptr())
            barrier()
            t0 = time.perf_counter()
            for _ in range(ssteps):
                ctx.raw("bn254_hash_to_g1" if grp == 1 else "bn254_hash_to_g2", hm3.data_ptr(), ho3.data_ptr(), n3, dstg, len(dstg), hout.data_ptr(), hst3.data_ptr())
            ems = max_over_ranks((time.perf_counter() - t0) * 1e3)
            assert bytes(hout[:rec * 4096].numpy()) == bytes(dout[:rec * 4096].cpu().numpy()) and int(hst3.min()) == 1
            check = None
            if rank == 0 and not args.no_cpu_baseline:
                from oracle import orc
                k3 = 2048 if grp == 1 else 512
                ref = (orc.hash_to_g1 if grp == 1 else orc.hash_to_g2)([bytes(r) for r in m3[:k3]], dstg, nthreads=orc.ncpu())
                assert ref == bytes(hout[:rec * k3].numpy()), "hash_to_g%d differs from the oracle" % grp
                check = "first %d points byte-equal to the oracle" % k3
            rec3["hash_g%d" % grp] = {"hashes_per_sec": n3 * world * ssteps / (ms * 1e-3), "ms_per_step": ms / ssteps,
                                      "e2e_hashes_per_sec": n3 * world * ssteps / (ems * 1e-3), "checked": check}
            del dout, dst3
        sub["c3"] = dict(rec3, config="hash_to_curve: 2^%d 32-byte messages per GPU to G1 and to G2 (expand_message_xmd SHA-256 + SvdW [+ cofactor clearing])" % args.c3_log2n,
                         steps=ssteps, scaling="weak")
        del dm3, do3
        torch.cuda.empty_cache()

    if "c4" in subs:
        # BASELINE config 4: 2^16 signatures in TOTAL as one multi-Miller product + a single final exponentiation, sharded over the ranks
        n4_total = 1 << 16
        lo4, hi4 = b.dist.shard_range(n4_total, world, rank)
        n4 = hi4 - lo4
        pk4, m4, s4, _ = b.scheme.synth_batch(ctx, n4, variant=0, first=(1 << 42) + lo4, n_keys=min(n4, N_KEYS), corrupt_every=0)
        agg4 = job_aggregate_signature(b.scheme.aggregate(ctx, s4, 0))
        blob4, off4 = b.pack_msgs(m4)
        dpk4, dbl4 = dev_bytes(pk4), dev_bytes(blob4)
        dof4 = torch.tensor(list(off4), dtype=torch.int64, device=dev)
        dag4 = dev_bytes(agg4) if rank == 0 else None
        dds4 = dev_bytes(b.DST_MIN_SIG)
        dok4 = torch.zeros(16, dtype=torch.uint8, device=dev)
        run = lambda: ctx.aggregate_verify_dist_dev(0, dpk4.data_ptr(), dbl4.data_ptr(), dof4.data_ptr(), n4, dag4.data_ptr() if rank == 0 else None,
                                                    dds4.data_ptr(), len(b.DST_MIN_SIG), 0, dok4.data_ptr())
        run(); ctx.sync()
        assert int(dok4[0].item()) == 1, "c4: a valid aggregate was rejected"
        barrier()
        ms, _ = timed_dev(run, ssteps)
        ms = max_over_ranks(ms)
        ctx.aggregate_verify_dist(0, pk4, (blob4, off4), agg4 if rank == 0 else None, b.DST_MIN_SIG)   # untimed
        barrier()
        t0 = time.perf_counter()
        for _ in range(ssteps):
            v4 = ctx.aggregate_verify_dist(0, pk4, (blob4, off4), agg4 if rank == 0 else None, b.DST_MIN_SIG)
        ems = max_over_ranks((time.perf_counter() - t0) * 1e3)
        assert v4 == 1
        # a tampered message on the LAST rank must flip the verdict on every rank (the exchange really carries the partials)
        m4b = list(m4)
        if rank == world - 1:
            m4b[-1] = b"tampered"
        assert ctx.aggregate_verify_dist(0, pk4, m4b, agg4 if rank == 0 else None, b.DST_MIN_SIG) == 0, "c4: a tampered message was accepted"
        check = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import orc
            k4 = 4096
            assert orc.aggregate_verify(0, pk4[:128 * k4], m4[:k4], orc.g1_sum(s4[:64 * k4]), b.DST_MIN_SIG, nthreads=orc.ncpu()) == 1
            check = "verdict equals the oracle's on the first 2^12 tuples"
        sub["c4"] = {"config": "aggregate verify: 2^16 signatures on distinct messages in total, sharded over %d rank(s); per rank hash -> Miller loops -> product "
                               "tree -> one 400-byte record; ONE ncclAllGather inside the library; one post-gather kernel (product, single final exponentiation, compare)" % world,
                     "signatures_total": n4_total, "signatures_per_sec": n4_total * ssteps / (ms * 1e-3), "ms_per_aggregate": ms / ssteps, "steps": ssteps,
                     "e2e_signatures_per_sec": n4_total * ssteps / (ems * 1e-3), "verdict": 1, "tampered_verdict": 0, "comm_nranks": ctx.comm_nranks if world > 1 else 1,
                     "scaling": "strong", "checked": check}
        del dpk4, dbl4, dof4

    if "c5" in subs:
        # BASELINE config 5: 2^24 tuples in TOTAL sharded over the ranks, through the host-buffer call (pinned buffers, chunk pipeline)
        n5_total = 1 << args.c5_log2n
        lo5, hi5 = b.dist.shard_range(n5_total, world, rank)
        n5 = hi5 - lo5
        tg0 = time.perf_counter()
        pk5, m5, of5, sg5, ex5 = b.scheme.synth_batch_np(ctx, n5, variant=0, first=(1 << 43) + lo5, n_keys=N_KEYS, corrupt_every=CORRUPT_EVERY)
        gen_s = time.perf_counter() - tg0
        hpk5, hm5 = torch.from_numpy(pk5.reshape(-1)).pin_memory(), torch.from_numpy(m5.reshape(-1)).pin_memory()
        hof5, hsg5 = torch.from_numpy(of5.view(np.int64)).pin_memory(), torch.from_numpy(sg5.reshape(-1)).pin_memory()
        hok5 = torch.zeros(n5, dtype=torch.uint8).pin_memory()
        wn = min(n5, 3 << 18)   # warm-up: three chunks, so that both staging sets (and the overlapped-tail buffers) exist at their final sizes
        ctx.raw("bn254_verify_batch", 0, hpk5.data_ptr(), hm5.data_ptr(), hof5.data_ptr(), hsg5.data_ptr(), wn, b.DST_MIN_SIG, len(b.DST_MIN_SIG), 0, hok5.data_ptr())
        barrier()
        t0 = time.perf_counter()
        ctx.raw("bn254_verify_batch", 0, hpk5.data_ptr(), hm5.data_ptr(), hof5.data_ptr(), hsg5.data_ptr(), n5, b.DST_MIN_SIG, len(b.DST_MIN_SIG), 0, hok5.data_ptr())
        exact_s = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(hok5.numpy(), ex5), "c5: accept bits differ from the construction"
        grp5 = 256
        ng5 = (n5 + grp5 - 1) // grp5
        want5 = np.minimum.reduceat(ex5, np.arange(0, n5, grp5))
        hgr5 = torch.zeros(ng5, dtype=torch.uint8).pin_memory()
        ctx.raw("bn254_verify_batch_randomized", 0, hpk5.data_ptr(), hm5.data_ptr(), hof5.data_ptr(), hsg5.data_ptr(), wn, b.DST_MIN_SIG, len(b.DST_MIN_SIG), 0,
                RBV_SEED, 64, grp5, hgr5.data_ptr())
        barrier()
        t0 = time.perf_counter()
        ctx.raw("bn254_verify_batch_randomized", 0, hpk5.data_ptr(), hm5.data_ptr(), hof5.data_ptr(), hsg5.data_ptr(), n5, b.DST_MIN_SIG, len(b.DST_MIN_SIG), 0,
                RBV_SEED, 64, grp5, hgr5.data_ptr())
        rand_s = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(hgr5.numpy(), want5), "c5: randomised group verdicts differ from the construction"
        check = None
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import orc
            lo = (1 << 18) - 128 if n5 > (1 << 18) else 0
            k5 = min(256, n5 - lo)
            ref = orc.verify_batch(0, pk5[lo:lo + k5].tobytes(), [bytes(r) for r in m5[lo:lo + k5]], sg5[lo:lo + k5].tobytes(), b.DST_MIN_SIG, nthreads=orc.ncpu())
            assert ref == bytes(hok5[lo:lo + k5].numpy()), "c5: accept bits differ from the oracle"
            check = "all accept bytes equal the construction; %d tuples straddling the first chunk boundary byte-equal to the oracle" % k5
        sub["c5"] = {"config": "batch verify 2^%d independent (pk, msg, sig) tuples in total (min-sig, %d keys, 1/%d corrupted) sharded over %d rank(s), "
                               "through bn254_verify_batch[_randomized] on pinned host buffers: chunks of one lane-pair wave (66 304 tuples = one 896-thread CTA per SM; randomised: 2^18), "
                               "upload / kernels / download overlapped on two staging sets" %
                               (args.c5_log2n, N_KEYS, CORRUPT_EVERY, world),
                     "tuples_total": n5_total, "tuples_per_rank": n5, "scaling": "strong",
                     "exact": {"seconds": exact_s, "verifications_per_sec": n5_total / exact_s},
                     "randomized": {"seconds": rand_s, "verifications_per_sec": n5_total / rand_s, "group": grp5, "rand_bits": 64},
                     "h2d_bytes_per_rank": int(pk5.nbytes + m5.nbytes + of5.nbytes + sg5.nbytes), "d2h_bytes_per_rank": int(n5),
                     "data_generation_s": gen_s, "checked": check}
        del hpk5, hm5, hof5, hsg5, pk5, m5, sg5

    c1 = None
    if "c1" in subs and rank == 0:
        # BASELINE config 1: sign + verify of ONE 32-byte message (latency).  The GPU library with n = 1, and the CPU port.
        sk = b.scheme.SecretKey(b.scheme.scalar_from_seed(b.scheme.SEED, 7))
        msg1 = [hashlib.sha256(b"config 1").digest()]
        pk1 = b.scheme.public_keys(ctx, [sk], 0)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            sig1 = b.scheme.sign(ctx, [sk], msg1, 0)
        t_sign = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            ok1 = ctx.verify_batch(0, pk1, msg1, sig1, b.DST_MIN_SIG)
        t_ver = (time.perf_counter() - t0) / reps
        assert ok1 == b"\x01"
        c1 = {"config": "sign + verify of one 32-byte message (latency of a lone call)", "gpu_sign_ms": t_sign * 1e3, "gpu_verify_ms": t_ver * 1e3}
        if not args.no_cpu_baseline:
            from oracle import orc
            t0 = time.perf_counter()
            for _ in range(reps):
                s_cpu = orc.g1_mul(orc.hash_to_g1(msg1, b.DST_MIN_SIG), sk.to_bytes())
            c1["cpu_sign_ms"] = (time.perf_counter() - t0) / reps * 1e3
            t0 = time.perf_counter()
            for _ in range(reps):
                ok_cpu = orc.verify_batch(0, pk1, msg1, s_cpu, b.DST_MIN_SIG, nthreads=1)
            c1["cpu_verify_ms"] = (time.perf_counter() - t0) / reps * 1e3
            assert s_cpu == sig1 and ok_cpu == b"\x01", "c1: GPU and CPU signatures differ"
            c1["note"] = "identical signature bytes; a lone pairing is a latency-bound chain, the GPU path is for batches"
        sub["c1"] = c1

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event durations from the timed region) ---------
    agg_t, by_name = {}, {}
    for name, items, ms in prof:          # launches of one kernel at different sizes are different roofline entries
        a = agg_t.setdefault((name, items), [0, 0.0, items])
        a[0] += 1
        a[1] += ms
        by_name[name] = by_name.get(name, 0.0) + ms
    kern_ms_total = sum(a[1] for a in agg_t.values())
    top = max(agg_t.items(), key=lambda kv: kv[1][1])
    top_name, (top_cnt, top_ms, top_items) = top[0][0], top[1]
    peak_name = max(peaks, key=peaks.get)
    peak = peaks[peak_name]
    execd = executed_ops()

    def base(name):
        return name[:-3] if name.endswith("_l2") else name

    def kernel_row(name, cnt, ms, items):
        lanes = 2 if name.endswith("_l2") else 1          # lane-pair kernels: two threads per item
        it = items // lanes
        per = ms / cnt * 1e-3
        row = {"ms_per_launch": ms / cnt, "items": it}
        if base(name) in KERNEL_M:
            row["canonical_M"] = KERNEL_M[base(name)]
            row["frac_canonical"] = KERNEL_M[base(name)] * M_IMAD * it / per / peak
        if base(name) in execd:
            row["executed_M"] = round(execd[base(name)] / M_IMAD, 1)
            row["frac_executed"] = execd[base(name)] * it / per / peak
        return row

    kernels = {}
    for (name, items), (cnt, ms, _) in sorted(agg_t.items(), key=lambda kv: -kv[1][1]):
        if ms / kern_ms_total >= 0.002:
            kernels["%s[%d]" % (name, items)] = kernel_row(name, cnt, ms, items)
    top_row = kernel_row(top_name, top_cnt, top_ms, top_items)
    top_ops = KERNEL_M.get(base(top_name), 0) * M_IMAD * top_row["items"]
    achieved = top_ops / (top_ms / top_cnt * 1e-3) if top_ms else 0.0
    # Where a kernel executes much less multiplier work than the canonical count assumes (hashing: Jacobi symbols and binary
    # GCDs on the ALU pipe instead of exponentiations), the canonical fraction is not a pipe utilisation and can pass 1:
    # `frac` is then the executed fraction, the canonical one is kept beside it as `frac_canonical`.
    frac = achieved / peak
    if "frac_executed" in top_row and top_row["frac_executed"] < 0.9 * frac:
        frac = top_row["frac_executed"]
        achieved = frac * peak
    whole_exec = sum(execd.get(base(nm), 0) * (it // (2 if nm.endswith("_l2") else 1)) * cnt for (nm, it), (cnt, ms, _) in agg_t.items())
    traffic, traffic_note = None, "no ncu capture of these kernel sources is committed"
    try:  # DRAM bytes of the dominant kernel from the committed ncu capture: same sources, same launch size only
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("source_sha16") != source_sha16():
            traffic_note = "profiles/traffic.json was captured from other kernel sources (%s): not quoted" % tr.get("source_sha16")
        elif tr["items_per_launch"] == top_row["items"] and top_name in tr["kernels"]:
            traffic = tr["kernels"][top_name]["dram_read_bytes"] + tr["kernels"][top_name]["dram_write_bytes"]
            traffic_note = "DRAM bytes per launch (ncu --set full, profiles/traffic.json, same kernel sources): per-thread stack write-back, not algorithmic bytes"
    except Exception:
        pass
    roofline = {
        "bound": "int32_imad", "kernel": top_name, "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TIMAD-op/s", "frac": frac,
        "frac_canonical": top_row.get("frac_canonical"), "frac_executed": top_row.get("frac_executed"),
        "traffic": traffic, "traffic_unit": traffic_note,
        "peak_source": "measured live by bn254_imad_peak (%s chains, all SMs); MEASURED_PEAKS.json has no INT32 figure" % peak_name,
        "peaks_measured_tops": {k: v / 1e12 for k, v in peaks.items()},
        "field_level_tops": field_peaks,
        "frac_note": "frac = canonical multiplier work of SURVEY.md 8(d) (items x M x 264 INT32 multiply-adds) / kernel time / peak; frac_executed counts the "
                     "work the kernel really issues (profiles/executed_ops.json, from the host build's op counter).  Per kernel under `kernels`.",
        "algorithmic_ops_per_launch": top_ops, "kernel_ms_per_launch": top_ms / top_cnt, "kernel_share_of_step": top_ms / kern_ms_total,
        "whole_step_frac_canonical": WORK_M[w] * M_IMAD * n * args.steps / (dev_ms * 1e-3) / peak,
        "whole_step_frac_executed": whole_exec / (prof_ms * 1e-3) / peak if whole_exec else None,
        "kernel_pass_ms_per_step": prof_ms / args.steps,
        "kernels": kernels,
        "kernels_ms_per_step": {k: v / args.steps for k, v in sorted(by_name.items(), key=lambda kv: -kv[1])},
        "hbm": {"algorithmic_bytes_per_step": in_bytes + out_bytes, "achieved_gbs": (in_bytes + out_bytes) / (dev_ms / args.steps * 1e-3) / 1e9,
                "peak_gbs": 6553.3, "note": "HBM non-binding"},
    }

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on the box's host cores ----------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        cores = orc.ncpu()
        m = min(n, args.cpu_sample)
        t0 = time.perf_counter()
        if w in ("verify", "verify_pk"):
            rpk, rsig = (128, 64) if variant == 0 else (64, 128)
            ref = orc.verify_batch(variant, pks[:m * rpk], msgs[:m], sigs[:m * rsig], dst, nthreads=cores)
            assert ref == expected[:m], "oracle and construction disagree"
        elif w == "verify_rand":
            m = max(args.group, m // args.group * args.group)
            ref = orc.verify_batch_randomized(0, pks[:m * 128], msgs[:m], sigs[:m * 64], dst, RBV_SEED, group=args.group, nthreads=cores)
            assert ref == expected[:len(ref)], "oracle rejected a valid group"
        elif w == "hash_g1":
            ref = orc.hash_to_g1(msgs[:m], dst, nthreads=cores)
            assert ref == bytes(d_out[:64 * m].cpu().numpy()), "hash_to_g1 differs from the oracle"
        elif w == "hash_g2":
            m = min(m, 4096)
            ref = orc.hash_to_g2(msgs[:m], dst, nthreads=cores)
            assert ref == bytes(d_out[:128 * m].cpu().numpy()), "hash_to_g2 differs from the oracle"
        elif w == "pairing":
            ref = orc.pairing_batch(g1[:64 * m], g2[:128 * m], nthreads=cores)
            assert ref == bytes(d_out[:384 * m].cpu().numpy()), "pairing differs from the oracle"
        else:
            assert orc.aggregate_verify(0, pks[:m * 128], msgs[:m], orc.g1_sum(sigs[:m * 64]), dst, nthreads=cores) == 1
        dt = time.perf_counter() - t0
        cpu = {"value": m / dt, "unit": UNIT[w], "cores": cores, "kind": "port",
               "sample": "first %d tuples of the same batch, oracle C port (4x64-bit Montgomery) on %d threads, %.1f s; results compared" % (m, cores, dt)}

    line = {
        "metric": METRIC[w], "value": value, "unit": UNIT[w], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": workload_config(w, args),
        "e2e": {"value": n * world * args.steps / (e2e_ms * 1e-3), "unit": UNIT[w], "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        # "pairing" = one Miller loop + one final exponentiation: the standalone bn254_pairing_batch rate
        "pairings_per_sec": sub["pairing"]["pairings_per_sec"] if "pairing" in sub else (value if w == "pairing" else None),
        "miller_loops_per_sec_inside_verify": value * 2 if w in ("verify", "verify_pk") else None,
    }
    line.update(sub)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
