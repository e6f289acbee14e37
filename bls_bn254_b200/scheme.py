"""BLS scheme layer over the GPU primitives: keys, sign, verify, aggregate, aggregate-verify.

The reference snapshot has no scheme layer (no SecretKey / PublicKey / Signature
anywhere under src/); north_star names it as the crate API, so it is defined here on top of the
`inner_types` operations exactly as SURVEY.md 8(a) row V states:
  min-sig: pk = sk*g2 (G2), sig = sk*H_G1(m) (G1), verify e(sig, -g2) * e(H_G1(m), pk) == 1
  min-pk : pk = sk*g1 (G1), sig = sk*H_G2(m) (G2), verify e(-g1, sig) * e(pk, H_G2(m)) == 1
All group work (hashing, scalar multiplication, sums, pairings) runs on the GPU through `Context`.
"""
import hashlib
import os
import struct

from .api import pack_msgs
from .api import CHECK_ON_CURVE, CHECK_SUBGROUP, DST_MIN_PK, DST_MIN_SIG, G1_GENERATOR, G2_GENERATOR, MIN_PK, MIN_SIG, R_ORDER

# Keys and signatures normally arrive from outside: the scheme-level predicates validate them by default (on the curve and
# in the order-r subgroup, inside the same GPU pass; +13 % time for min-sig).  flags=0 is for inputs validated earlier.
VALIDATE = CHECK_ON_CURVE | CHECK_SUBGROUP

SEED = 0xB15B254


def _dst(variant):
    return DST_MIN_SIG if variant == MIN_SIG else DST_MIN_PK


def scalar_from_seed(seed, index):
    """Deterministic non-zero scalar mod r for (seed, index): SHA-256 counter stream."""
    v = 0
    ctr = 0
    while v == 0:
        h = hashlib.sha256(struct.pack("<QQQ", seed, index, ctr)).digest() + hashlib.sha256(struct.pack("<QQQ", seed ^ 0x5EED, index, ctr)).digest()
        v = int.from_bytes(h, "big") % R_ORDER
        ctr += 1
    return v


class SecretKey:
    def __init__(self, value):
        if not 0 < value < R_ORDER:
            raise ValueError("secret key out of range")
        self.value = value

    def to_bytes(self):
        return self.value.to_bytes(32, "big")


def public_keys(ctx, sks, variant=MIN_SIG):
    """pk_i = sk_i * generator (wire bytes, concatenated)."""
    sc = b"".join(s.to_bytes() for s in sks)
    if variant == MIN_SIG:
        return ctx.g2_mul(G2_GENERATOR * len(sks), sc)
    return ctx.g1_mul(G1_GENERATOR * len(sks), sc)


def sign(ctx, sks, msgs, variant=MIN_SIG, dst=None):
    """sig_i = sk_i * H(m_i) (wire bytes, concatenated)."""
    dst = _dst(variant) if dst is None else dst
    sc = b"".join(s.to_bytes() for s in sks)
    if variant == MIN_SIG:
        return ctx.g1_mul(ctx.hash_to_g1(msgs, dst), sc)
    return ctx.g2_mul(ctx.hash_to_g2(msgs, dst), sc)


def verify(ctx, pks, msgs, sigs, variant=MIN_SIG, dst=None, flags=VALIDATE):
    """n accept bytes (0/1) for n independent (pk, msg, sig) tuples.  Identity keys / signatures never verify."""
    return ctx.verify_batch(variant, pks, msgs, sigs, _dst(variant) if dst is None else dst, flags)


GLV_LAMBDA = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd   # eigenvalue of (x, y) -> (beta x, y); oracle/pymodel.py::glv_constants


def rbv_scalar(seed, index, rand_bits=64):
    """The per-tuple scalar of the randomised batch test (include/bls_bn254_b200.h): a + b lambda mod r."""
    dg = hashlib.sha256(bytes(seed) + struct.pack("<Q", index)).digest()
    h = rand_bits // 16
    a, b = int.from_bytes(dg[:h], "big"), int.from_bytes(dg[h:2 * h], "big")
    if a == 0 and b == 0:
        a = 1
    return (a + b * GLV_LAMBDA) % R_ORDER


def verify_randomized(ctx, pks, msgs, sigs, variant=MIN_SIG, dst=None, flags=VALIDATE, group=256, rand_bits=64, seed=None):
    """n accept bytes like verify(), computed with one final exponentiation per GROUP of tuples: groups that
    pass the small-exponent test are accepted wholesale (error probability <= 2^-rand_bits per group), the
    tuples of a rejected group are re-verified one by one with the exact predicate.  `seed` defaults to 32
    fresh bytes from the OS CSPRNG; pass one only for reproducible tests.
    The error bound holds for keys and signatures in the order-r subgroups ([a + b lambda]P and the group sums
    mean nothing elsewhere, and G2 has a cofactor with small factors): the default flags check that; pass
    flags=0 only for inputs that were validated before."""
    dst = _dst(variant) if dst is None else dst
    seed = os.urandom(32) if seed is None else seed
    blob, off = pack_msgs(msgs) if not isinstance(msgs, tuple) else msgs
    n = len(off) - 1
    if n == 0:
        return b""
    groups = ctx.verify_batch_randomized(variant, pks, (blob, off), sigs, dst, seed, group=group, rand_bits=rand_bits, flags=flags)
    out = bytearray(n)
    rpk, rsig = (128, 64) if variant == MIN_SIG else (64, 128)
    redo = []
    for g, okg in enumerate(groups):
        lo, hi = g * group, min(n, (g + 1) * group)
        if okg:
            out[lo:hi] = b"\x01" * (hi - lo)
        else:
            redo.extend(range(lo, hi))
    if redo:
        blob_b = bytes(blob)
        sub_msgs = [blob_b[off[i]:off[i + 1]] for i in redo]
        sub_pks = b"".join(pks[i * rpk:(i + 1) * rpk] for i in redo)
        sub_sigs = b"".join(sigs[i * rsig:(i + 1) * rsig] for i in redo)
        exact = ctx.verify_batch(variant, sub_pks, sub_msgs, sub_sigs, dst, flags)
        for i, v in zip(redo, exact):
            out[i] = v
    return bytes(out)


def verify_same_key(ctx, pks, msg_counts, msgs, sigs, variant=MIN_SIG, dst=None, flags=VALIDATE, rand_bits=64, seed=None):
    """Many messages under few keys: key k signed the next msg_counts[k] messages (tuples sorted by key).  Returns n accept
    bytes like verify(): every key first gets ONE randomised pairing check over all its tuples
    (bn254_verify_same_key_randomized: error probability <= 2^-rand_bits per key); the tuples of a key that passes are accepted
    wholesale, those of a rejected key are re-verified one by one with the exact predicate.  `seed` and `flags` as in
    verify_randomized()."""
    dst = _dst(variant) if dst is None else dst
    seed = os.urandom(32) if seed is None else seed
    blob, off = pack_msgs(msgs) if not isinstance(msgs, tuple) else msgs
    n = len(off) - 1
    if n == 0:
        return b""
    keys_ok = ctx.verify_same_key_randomized(variant, pks, msg_counts, (blob, off), sigs, dst, seed, rand_bits=rand_bits, flags=flags)
    out = bytearray(n)
    rpk, rsig = (128, 64) if variant == MIN_SIG else (64, 128)
    redo, redo_key, lo = [], [], 0
    for k, c in enumerate(msg_counts):
        if keys_ok[k]:
            out[lo:lo + c] = b"\x01" * c
        else:
            redo.extend(range(lo, lo + c))
            redo_key.extend([k] * c)
        lo += c
    if redo:
        blob_b = bytes(blob)
        sub_msgs = [blob_b[off[i]:off[i + 1]] for i in redo]
        sub_pks = b"".join(pks[k * rpk:(k + 1) * rpk] for k in redo_key)
        sub_sigs = b"".join(sigs[i * rsig:(i + 1) * rsig] for i in redo)
        for i, v in zip(redo, ctx.verify_batch(variant, sub_pks, sub_msgs, sub_sigs, dst, flags)):
            out[i] = v
    return bytes(out)


def aggregate(ctx, sigs, variant=MIN_SIG):
    """Sum of signatures (one wire-format point)."""
    return ctx.g1_sum(sigs) if variant == MIN_SIG else ctx.g2_sum(sigs)


def aggregate_verify(ctx, pks, msgs, agg_sig, variant=MIN_SIG, dst=None, flags=VALIDATE):
    """Distinct-message aggregate verification: one multi-Miller loop, one final exponentiation."""
    return ctx.aggregate_verify(variant, pks, msgs, agg_sig, _dst(variant) if dst is None else dst, flags)


def synth_batch(ctx, n, variant=MIN_SIG, seed=SEED, first=0, msg_len=32, corrupt_every=64, n_keys=None):
    """Synthetic workload of SURVEY.md 8(d): random keys (n_keys distinct ones, default one per
    tuple), distinct `msg_len`-byte messages (first 8 bytes = index), signatures made on the GPU;
    every `corrupt_every`-th tuple is broken (bit flip in the message, somebody else's key, or a valid
    signature of another tuple) so both verdicts occur.
    Returns (pks, msgs list, sigs, expected accept bytes)."""
    nk = n if not n_keys else min(n_keys, n)
    uniq = [SecretKey(scalar_from_seed(seed, first + i if nk == n else i)) for i in range(nk)]
    sks = [uniq[i % nk] for i in range(n)]
    msgs = []
    for i in range(first, first + n):
        body = hashlib.sha256(struct.pack("<QQ", seed, i)).digest()
        stream = struct.pack("<Q", i) + body * (msg_len // 32 + 1)
        msgs.append(stream[:msg_len] if msg_len >= 8 else body[:msg_len])
    rpk, rsig = (128, 64) if variant == MIN_SIG else (64, 128)
    upk = public_keys(ctx, uniq, variant)
    pks = bytearray(b"".join(upk[(i % nk) * rpk:(i % nk + 1) * rpk] for i in range(n)) if nk != n else upk)
    sigs = bytearray(sign(ctx, sks, msgs, variant))
    expected = bytearray(b"\x01" * n)
    if corrupt_every and n > 1:
        for j in range(corrupt_every // 2, n, corrupt_every):
            kind = (j // corrupt_every) % 3
            o = (j + 1) % n
            if kind == 0 and msg_len > 0:      # bit flip in the message
                b = bytearray(msgs[j])
                b[-1] ^= 1
                msgs[j] = bytes(b)
            elif kind == 1 and sks[o].value != sks[j].value:   # somebody else's key
                pks[j * rpk:(j + 1) * rpk] = pks[o * rpk:(o + 1) * rpk]
            else:                              # a valid signature of another tuple
                sigs[j * rsig:(j + 1) * rsig] = sigs[o * rsig:(o + 1) * rsig]
            expected[j] = 0
    return bytes(pks), msgs, bytes(sigs), bytes(expected)


def _splitmix64(x):
    """numpy uint64 array -> SplitMix64 output (wrapping arithmetic)"""
    import numpy as np
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_messages_np(n, first=0, seed=SEED):
    """n distinct 32-byte messages as an (n, 32) uint8 array: bytes 0..7 = the tuple's global index (little-endian, which
    makes them distinct), bytes 8..31 = SplitMix64(seed, index) output.  Vectorised: 2^24 messages take about a second."""
    import numpy as np
    idx = np.arange(first, first + n, dtype=np.uint64)
    cols = [idx]
    with np.errstate(over="ignore"):
        for k in range(3):
            cols.append(_splitmix64(idx * np.uint64(3) + np.uint64(k) + np.uint64(seed) * np.uint64(0x100000001B3)))
    return np.ascontiguousarray(np.stack(cols, axis=1).astype("<u8")).view(np.uint8).reshape(n, 32)


def synth_batch_np(ctx, n, variant=MIN_SIG, seed=SEED, first=0, n_keys=4096, corrupt_every=64):
    """The synthetic workload of SURVEY.md 8(d) for BIG batches (BASELINE config 5: 2^24 tuples): the same construction as
    synth_batch -- n_keys distinct random keys used round-robin, distinct 32-byte messages, signatures made on the GPU,
    every corrupt_every-th tuple broken in one of three ways -- with every per-tuple step vectorised in numpy.
    Returns numpy arrays (pks [n, rpk], msgs [n, 32], offsets [n + 1] u64, sigs [n, rsig], expected [n])."""
    import numpy as np
    nk = min(n_keys, n)
    uniq = [SecretKey(scalar_from_seed(seed, i)) for i in range(nk)]
    rpk, rsig = (128, 64) if variant == MIN_SIG else (64, 128)
    upk = np.frombuffer(public_keys(ctx, uniq, variant), dtype=np.uint8).reshape(nk, rpk)
    usc = np.frombuffer(b"".join(s.to_bytes() for s in uniq), dtype=np.uint8).reshape(nk, 32)
    key_of = (np.arange(first, first + n, dtype=np.int64) % nk)
    pks = upk[key_of].copy()
    msgs = synth_messages_np(n, first, seed)
    off = np.arange(0, 32 * (n + 1), 32, dtype=np.uint64)
    sc = np.ascontiguousarray(usc[key_of]).reshape(-1)     # 1-D uint8 views: len() is the byte count the wrappers check
    dst = _dst(variant)
    if variant == MIN_SIG:
        sig_b = ctx.g1_mul(ctx.hash_to_g1((msgs.reshape(-1), off), dst), sc)
    else:
        sig_b = ctx.g2_mul(ctx.hash_to_g2((msgs.reshape(-1), off), dst), sc)
    sigs = np.frombuffer(sig_b, dtype=np.uint8).reshape(n, rsig).copy()
    expected = np.ones(n, dtype=np.uint8)
    if corrupt_every and n > 1:
        j = np.arange(corrupt_every // 2, n, corrupt_every, dtype=np.int64)
        kind = (j // corrupt_every) % 3
        o = (j + 1) % n
        same_key = key_of[o] == key_of[j]
        j0 = j[kind == 0]
        msgs[j0, 31] ^= 1                                   # bit flip in the message
        m1 = (kind == 1) & ~same_key
        pks[j[m1]] = pks[o[m1]]                             # somebody else's key
        m2 = (kind == 2) | ((kind == 1) & same_key)
        sigs[j[m2]] = sigs[o[m2]]                           # a valid signature of another tuple
        expected[j] = 0
    return pks, msgs, off, sigs, expected


def synth_same_key_np(ctx, n, n_keys=64, variant=MIN_SIG, seed=SEED, first=0, bad_keys=()):
    """Workload of the same-key path: n_keys keys, key k owns the next n // n_keys (+1 for the first n % n_keys keys) tuples --
    distinct 32-byte messages, signatures made on the GPU; every key listed in bad_keys gets ONE broken tuple (a valid signature
    of its neighbour).  Returns numpy arrays (pks [n_keys, rpk], key_offsets [n_keys + 1] u64, msgs [n, 32], offsets [n + 1] u64,
    sigs [n, rsig], expected [n_keys])."""
    import numpy as np
    rpk, rsig = (128, 64) if variant == MIN_SIG else (64, 128)
    uniq = [SecretKey(scalar_from_seed(seed, (1 << 20) + i)) for i in range(n_keys)]
    upk = np.frombuffer(public_keys(ctx, uniq, variant), dtype=np.uint8).reshape(n_keys, rpk).copy()
    usc = np.frombuffer(b"".join(s.to_bytes() for s in uniq), dtype=np.uint8).reshape(n_keys, 32)
    counts = np.full(n_keys, n // n_keys, dtype=np.uint64)
    counts[:n % n_keys] += np.uint64(1)
    koff = np.zeros(n_keys + 1, dtype=np.uint64)
    koff[1:] = np.cumsum(counts)
    key_of = np.repeat(np.arange(n_keys, dtype=np.int64), counts.astype(np.int64))
    msgs = synth_messages_np(n, first, seed)
    off = np.arange(0, 32 * (n + 1), 32, dtype=np.uint64)
    sc = np.ascontiguousarray(usc[key_of]).reshape(-1)
    dst = _dst(variant)
    hashed = (ctx.hash_to_g1 if variant == MIN_SIG else ctx.hash_to_g2)((msgs.reshape(-1), off), dst)
    sigs = np.frombuffer((ctx.g1_mul if variant == MIN_SIG else ctx.g2_mul)(hashed, sc), dtype=np.uint8).reshape(n, rsig).copy()
    expected = (counts > 0).astype(np.uint8)
    for k in bad_keys:
        if counts[k] >= 2:
            j = int(koff[k]) + int(counts[k]) // 2
            sigs[j] = sigs[j - 1]
            expected[k] = 0
    return upk, koff, msgs, off, sigs, expected
