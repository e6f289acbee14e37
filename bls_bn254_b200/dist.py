"""One-process-per-GPU sharding of the path over torch.distributed.

Independent verifications / hashes / pairings shard by index with NO data-path collective: every
rank runs its contiguous slice [n*rank/world, n*(rank+1)/world) on its own GPU.  Aggregate
verification has exactly one exchange step: each rank reduces its Miller loops to one Fp12 partial
product, the partials are all-gathered, and one final exponentiation decides.  On GPU ranks the
exchange lives INSIDE the library (`init_comm` hands it an NCCL communicator; the product tree's last
pass writes the send record, `ncclAllGather` runs on the library stream, one kernel finishes --
`Context.aggregate_verify_dist`); torch.distributed only ships the 128-byte NCCL unique id once.  A
ctx without that entry point (the CPU stand-in of the gloo tests) takes the same steps through
`gather_partials` with torch.distributed doing the all-gather.
"""
import hashlib
import struct

import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    """Contiguous slice of n items owned by `rank` (same split as the C ABI's multi-device ctx)."""
    return n * rank // world, n * (rank + 1) // world


def shard_tuples(n, world, rank, records):
    """records: list of (buffer, record_size) -> the rank's slices (messages: pass the list itself with size None)."""
    lo, hi = shard_range(n, world, rank)
    out = []
    for buf, rec in records:
        out.append(buf[lo:hi] if rec is None else buf[lo * rec:hi * rec])
    return out


def init_comm(ctx, device=None, group=None):
    """Give `ctx` (one device, this rank's) the job's NCCL communicator: rank 0 draws the unique id in the library and
    torch.distributed broadcasts its 128 bytes.  No-op for a single-rank job or a ctx that is already initialised."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1 or getattr(ctx, "comm_nranks", 1) == world:
        return world
    rank = dist.get_rank(group)
    uid = ctx.comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init_rank(bytes(t.cpu().numpy()), world, rank)
    return world


def gather_partials(partial, bad, device=None, group=None):
    """all-gather of the (384 + 1)-byte partial records -> (concatenated partials, list of bad flags)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return partial, [bad]
    t = torch.tensor(list(partial) + [bad], dtype=torch.uint8, device=device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    rows = [bytes(o.cpu().numpy()) for o in outs]
    return b"".join(r[:384] for r in rows), [r[384] for r in rows]


VALIDATE = 3   # CHECK_ON_CURVE | CHECK_SUBGROUP: the default for keys and signatures that arrive from outside


def aggregate_verify_sharded(ctx, variant, pks_shard, msgs_shard, agg_sig, dst, flags=VALIDATE, device=None, group=None):
    """Every rank passes ITS shard of (pk_i, m_i); `agg_sig` is used on rank 0 only.  Returns the
    verdict (0/1) on every rank.  `ctx` needs aggregate_partial / aggregate_finish (bls_bn254_b200.Context)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if hasattr(ctx, "aggregate_verify_dist"):   # the library's own exchange (NCCL all-gather + one post-gather kernel)
        init_comm(ctx, device=device, group=group)
        return ctx.aggregate_verify_dist(variant, pks_shard, msgs_shard, agg_sig if rank == 0 else None, dst, flags)
    part, bad = ctx.aggregate_partial(variant, pks_shard, msgs_shard, agg_sig if rank == 0 else None, dst, flags)
    parts, bads = gather_partials(part, bad, device=device, group=group)
    return ctx.aggregate_finish(parts, bads)


def verify_sharded(ctx, variant, pks_shard, msgs_shard, sigs_shard, dst, flags=VALIDATE, gather=False, device=None, group=None):
    """Every rank verifies its own shard (no collective).  gather=True additionally all-gathers the
    accept bytes so each rank holds the whole job's verdicts (shards must be equally long)."""
    ok = ctx.verify_batch(variant, pks_shard, msgs_shard, sigs_shard, dst, flags)
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ok
    t = torch.tensor(list(ok), dtype=torch.uint8, device=device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t, group=group)
    return b"".join(bytes(o.cpu().numpy()) for o in outs)


def verify_randomized_sharded(ctx, variant, pks_shard, msgs_shard, sigs_shard, dst, seed, group_size=256, rand_bits=64, flags=VALIDATE,
                              device=None, group=None):
    """Randomised batch verification of a sharded job: every rank runs the small-exponent test on ITS shard (groups never
    span ranks, so there is no data-path collective) with a per-rank seed derived from the common `seed`, then one
    all-reduce (MIN) of a single byte tells every rank whether the whole job was accepted.
    The 2^-rand_bits error bound of the small-exponent test needs keys and signatures in the order-r subgroups, hence the
    default flags (on-curve + subgroup checks); pass flags=0 only for inputs validated earlier.
    Returns (this rank's group verdicts, whole-job verdict 0/1)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rank_seed = hashlib.sha256(bytes(seed) + struct.pack("<I", rank)).digest()
    local = ctx.verify_batch_randomized(variant, pks_shard, msgs_shard, sigs_shard, dst, rank_seed, group=group_size, rand_bits=rand_bits, flags=flags)
    ok = int(all(local))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([ok], dtype=torch.uint8, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        ok = int(t.item())
    return local, ok


def shard_keys(msg_counts, world, rank):
    """Same-key jobs shard by KEY: rank r owns the contiguous keys [nk r / world, nk (r + 1) / world) and with them the tuples
    [sum(counts[:k_lo]), sum(counts[:k_hi])).  Returns (k_lo, k_hi, t_lo, t_hi)."""
    k_lo, k_hi = shard_range(len(msg_counts), world, rank)
    t_lo = sum(msg_counts[:k_lo])
    return k_lo, k_hi, t_lo, t_lo + sum(msg_counts[k_lo:k_hi])


def verify_same_key_sharded(ctx, variant, pks_shard, msg_counts_shard, msgs_shard, sigs_shard, dst, seed, rand_bits=64, flags=VALIDATE,
                            gather=False, device=None, group=None):
    """Same-key randomised verification of a sharded job: every rank checks ITS keys (a key and all its tuples live on one
    rank, so there is no data-path collective) with a per-rank seed derived from the common `seed`; one all-reduce (MIN) of a
    single byte tells every rank whether the whole job was accepted, and gather=True additionally all-gathers the per-key
    verdicts (padded to the longest shard) so that every rank holds all of them in key order.
    Returns (this rank's per-key verdicts or, with gather, the whole job's; whole-job verdict 0/1)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank_seed = hashlib.sha256(bytes(seed) + struct.pack("<I", rank)).digest()
    local = ctx.verify_same_key_randomized(variant, pks_shard, msg_counts_shard, msgs_shard, sigs_shard, dst, rank_seed, rand_bits=rand_bits, flags=flags)
    ok = int(all(local)) if len(local) else 1
    if world == 1:
        return local, ok
    t = torch.tensor([ok], dtype=torch.uint8, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    ok = int(t.item())
    if not gather:
        return local, ok
    n = torch.tensor([len(local)], dtype=torch.int64, device=device)
    lens = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(lens, n, group=group)
    longest = max(int(x.item()) for x in lens)
    pad = torch.zeros(max(1, longest), dtype=torch.uint8, device=device)
    if len(local):
        pad[:len(local)] = torch.tensor(list(local), dtype=torch.uint8, device=device)
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return b"".join(bytes(o.cpu().numpy())[:int(l.item())] for o, l in zip(outs, lens)), ok
