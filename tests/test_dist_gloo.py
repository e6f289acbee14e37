"""world_size-2 gloo test of the N>1 host path (bls_bn254_b200/dist.py): shard ranges, the single
all-gather of Fp12 partials for aggregate verification, gathered accept bytes, the all-reduced verdict of the
randomised batch verification, and the same-key path sharded by key.  No GPU here, so the
per-rank compute is a stand-in built on the CPU oracle (tests may use it); the plumbing under test
(sharding, record layout, collective, finish) is the code the GPU ranks run."""
import os
import pathlib
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parent.parent
DST = b"BLS_SIG_BN254G1_XMD:SHA-256_SVDW_RO_NUL_"
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


class OracleCtx:
    """aggregate_partial / aggregate_finish / verify_batch with the C-ABI semantics, on the oracle"""

    def __init__(self):
        from oracle import orc
        self.o = orc
        g = orc.g2_generator()
        neg = lambda b: ((P - int.from_bytes(b, "big")) % P).to_bytes(32, "big")
        self.neg_g2 = g[:64] + neg(g[64:96]) + neg(g[96:])

    def aggregate_partial(self, variant, pks, msgs, agg_sig, dst, flags=0):
        h = self.o.hash_to_g1(msgs, dst)
        g1, g2 = h, pks
        if agg_sig is not None:
            g1, g2 = g1 + agg_sig, g2 + self.neg_g2
        try:
            return self.o.multi_miller_loop(g1, g2), 0
        except ValueError:
            return bytes(384), 1

    def aggregate_finish(self, parts, bads):
        if any(bads):
            return 0
        one = (1).to_bytes(32, "big") + bytes(352)
        return int(self.o.final_exponentiation(self.o.fp12_product(parts)) == one)

    # a fake communicator so that bd.init_comm (rank 0 draws the id, torch.distributed ships it) is exercised on gloo
    comm = None

    @property
    def comm_nranks(self):
        return self.comm[1] if self.comm else 1

    def comm_unique_id(self):
        return bytes((7 * i + 3) & 0xff for i in range(128))

    def comm_init_rank(self, uid, nranks, rank):
        assert self.comm is None and len(uid) == 128
        self.comm = (bytes(uid), nranks, rank)

    def verify_batch(self, variant, pks, msgs, sigs, dst, flags=0):
        return self.o.verify_batch(variant, pks, msgs, sigs, dst, flags)

    def verify_batch_randomized(self, variant, pks, msgs, sigs, dst, seed, group=256, rand_bits=64, flags=0):
        return self.o.verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, group=group, rand_bits=rand_bits, flags=flags)

    def verify_same_key_randomized(self, variant, pks, msg_counts, msgs, sigs, dst, seed, rand_bits=64, flags=0):
        return self.o.verify_same_key_randomized(variant, pks, msg_counts, msgs, sigs, dst, seed, rand_bits=rand_bits, flags=flags)


def _make(n):
    import hashlib
    from oracle import orc
    sks = [int.from_bytes(hashlib.sha256(b"k%d" % i).digest(), "big") % R for i in range(n)]
    msgs = [b"m%d" % i for i in range(n)]
    sc = b"".join(s.to_bytes(32, "big") for s in sks)
    pks = orc.g2_mul(orc.g2_generator() * n, sc)
    sigs = orc.g1_mul(orc.hash_to_g1(msgs, DST), sc)
    return pks, msgs, sigs


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import hashlib
    import torch.distributed as dist
    from bls_bn254_b200 import dist as bd
    from oracle import orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 11
        pks, msgs, sigs = _make(n)
        agg = orc.g1_sum(sigs)
        ctx = OracleCtx()
        lo, hi = bd.shard_range(n, world, rank)
        p, m, s = bd.shard_tuples(n, world, rank, [(pks, 128), (msgs, None), (sigs, 64)])
        assert len(m) == hi - lo and len(p) == 128 * (hi - lo)
        assert bd.init_comm(ctx) == world and bd.init_comm(ctx) == world      # second call: already initialised
        assert ctx.comm == (OracleCtx().comm_unique_id(), world, rank)
        ok = bd.aggregate_verify_sharded(ctx, 0, p, m, agg, DST)
        bad_msgs = list(m)
        if rank == 1:
            bad_msgs[0] = b"tampered"
        ok_bad = bd.aggregate_verify_sharded(ctx, 0, p, bad_msgs, agg, DST)
        badkey = (b"\xff" * 128 + p[128:]) if rank == 1 else p
        ok_badkey = bd.aggregate_verify_sharded(ctx, 0, badkey, m, agg, DST)
        # independent verification: equal shards, gathered accept bytes
        n2 = 8
        lo2, hi2 = bd.shard_range(n2, world, rank)
        sig2 = bytearray(sigs[:64 * n2])
        sig2[64 * 5:64 * 6] = sigs[64 * 6:64 * 7]
        got = bd.verify_sharded(ctx, 0, pks[128 * lo2:128 * hi2], msgs[lo2:hi2], bytes(sig2[64 * lo2:64 * hi2]), DST, gather=True)
        # randomised batch verification: per-rank groups, one all-reduce of the verdict
        seed = bytes(range(32))
        rnd_good = bd.verify_randomized_sharded(ctx, 0, p, m, s, DST, seed, group_size=3)
        rnd_bad = bd.verify_randomized_sharded(ctx, 0, p, bad_msgs, s, DST, seed, group_size=3)
        # same-key grouping: 5 keys with 3, 0, 2, 4, 1 messages; keys (with their tuples) are what the ranks share out
        counts = [3, 0, 2, 4, 1]
        nk, nt = len(counts), sum(counts)
        key_of = [k for k, c in enumerate(counts) for _ in range(c)]
        kmsgs = [b"stream %d" % i for i in range(nt)]
        ksc = b"".join(hashlib.sha256(b"k%d" % key_of[i]).digest() for i in range(nt))
        ksc = b"".join((int.from_bytes(ksc[32 * i:32 * i + 32], "big") % R).to_bytes(32, "big") for i in range(nt))
        ksigs = bytearray(orc.g1_mul(orc.hash_to_g1(kmsgs, DST), ksc))
        ksigs[64 * 6:64 * 7] = ksigs[64 * 7:64 * 8]                      # key 3's second tuple carries its third signature
        k_lo, k_hi, t_lo, t_hi = bd.shard_keys(counts, world, rank)
        sk = bd.verify_same_key_sharded(ctx, 0, pks[128 * k_lo:128 * k_hi], counts[k_lo:k_hi], kmsgs[t_lo:t_hi], bytes(ksigs[64 * t_lo:64 * t_hi]), DST,
                                        seed, flags=0, gather=True)
        q.put((rank, ok, ok_bad, ok_badkey, got, rnd_good, rnd_bad, sk))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, ok_bad, ok_badkey, got, rnd_good, rnd_bad, sk in res:
        # keys 0, 1 on rank 0 and 2, 3, 4 on rank 1: the key without tuples and the key with a swapped signature reject
        assert sk == (b"\x01\x00\x01\x00\x01", 0)
        assert ok == 1 and ok_bad == 0 and ok_badkey == 0
        assert got == b"\x01\x01\x01\x01\x01\x00\x01\x01"
        # 11 tuples over 2 ranks = 5 + 6: two groups of <= 3 per rank; rank 1's first message was tampered with
        assert rnd_good == (b"\x01\x01", 1)
        assert rnd_bad == ((b"\x00\x01", 0) if rank == 1 else (b"\x01\x01", 0))


def test_shard_range_covers_everything():
    from bls_bn254_b200 import dist as bd
    for n in (0, 1, 7, 64, 1 << 16, (1 << 24) + 3):
        for world in (1, 2, 4, 8):
            cuts = [bd.shard_range(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
