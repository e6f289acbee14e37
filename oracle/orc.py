"""ctypes binding of the C oracle (oracle/liborc_bn254.so).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm."""
import ctypes as C
import os
import pathlib
import subprocess

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


def build(force=False):
    so = _HERE / "liborc_bn254.so"
    srcs = [_HERE / "bn254_oracle.c", _HERE / "orc_curve_tmpl.h", _HERE / "orc_constants.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-s", "-C", str(_HERE)])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_mul_count.restype = C.c_uint64
    return _LIB


def ncpu():
    return len(os.sched_getaffinity(0))


def pack_msgs(msgs):
    """list[bytes] -> (blob bytes, (n+1) u64 offsets array)"""
    off = (C.c_uint64 * (len(msgs) + 1))()
    pos = 0
    for i, m in enumerate(msgs):
        off[i] = pos
        pos += len(m)
    off[len(msgs)] = pos
    return b"".join(msgs), off


def _buf(n):
    return C.create_string_buffer(n)


def hash_to_g1(msgs, dst, count=2, nthreads=1, want_status=False):
    blob, off = pack_msgs(msgs)
    out, st = _buf(64 * len(msgs)), _buf(len(msgs))
    lib().orc_hash_to_g1(blob, off, C.c_size_t(len(msgs)), dst, C.c_size_t(len(dst)), count, out, st, nthreads)
    return (out.raw, st.raw) if want_status else out.raw


def hash_to_g2(msgs, dst, count=2, nthreads=1, want_status=False):
    blob, off = pack_msgs(msgs)
    out, st = _buf(128 * len(msgs)), _buf(len(msgs))
    lib().orc_hash_to_g2(blob, off, C.c_size_t(len(msgs)), dst, C.c_size_t(len(dst)), count, out, st, nthreads)
    return (out.raw, st.raw) if want_status else out.raw


def hash_to_field_fp(msgs, dst, count):
    blob, off = pack_msgs(msgs)
    out = _buf(32 * count * len(msgs))
    lib().orc_hash_to_field_fp(blob, off, C.c_size_t(len(msgs)), dst, C.c_size_t(len(dst)), count, out)
    return out.raw


def expand_message_xmd(msg, dst, n):
    out = _buf(n)
    rc = lib().orc_expand_message_xmd(msg, C.c_size_t(len(msg)), dst, C.c_size_t(len(dst)), out, C.c_size_t(n))
    assert rc == 0
    return out.raw


def svdw_g1(u_bytes):
    n = len(u_bytes) // 32
    out = _buf(64 * n)
    lib().orc_svdw_g1(u_bytes, C.c_size_t(n), out)
    return out.raw


def svdw_g2(u_bytes):
    n = len(u_bytes) // 64
    out = _buf(128 * n)
    lib().orc_svdw_g2(u_bytes, C.c_size_t(n), out)
    return out.raw


def pairing_batch(g1, g2, nthreads=1, want_status=False):
    n = len(g1) // 64
    out, st = _buf(384 * n), _buf(n)
    lib().orc_pairing_batch(g1, g2, C.c_size_t(n), out, st, nthreads)
    return (out.raw, st.raw) if want_status else out.raw


def miller_loop_batch(g1, g2, nthreads=1):
    n = len(g1) // 64
    out, st = _buf(384 * n), _buf(n)
    lib().orc_miller_loop_batch(g1, g2, C.c_size_t(n), out, st, nthreads)
    return out.raw


def g2_prepare(pts):
    """88 x 192 bytes of line coefficients per point (G2Prepared), status byte per point"""
    n = len(pts) // 128
    out, st = _buf(88 * 192 * n), _buf(n)
    lib().orc_g2_prepare(pts, C.c_size_t(n), out, st)
    return out.raw, st.raw


def multi_miller_loop(g1, g2):
    n = len(g1) // 64
    out = _buf(384)
    rc = lib().orc_multi_miller_loop(g1, g2, C.c_size_t(n), out)
    if rc:
        raise ValueError("orc_multi_miller_loop rc=%d" % rc)
    return out.raw


def final_exponentiation(ml, nthreads=1):
    n = len(ml) // 384
    out, st = _buf(384 * n), _buf(n)
    lib().orc_final_exponentiation(ml, C.c_size_t(n), out, st, nthreads)
    return out.raw


def fp12_product(elems):
    out = _buf(384)
    rc = lib().orc_fp12_product(elems, C.c_size_t(len(elems) // 384), out)
    assert rc == 0
    return out.raw


def verify_batch(variant, pks, msgs, sigs, dst, flags=0, nthreads=1):
    blob, off = pack_msgs(msgs)
    ok = _buf(len(msgs))
    lib().orc_verify_batch(variant, pks, blob, off, sigs, C.c_size_t(len(msgs)), dst, C.c_size_t(len(dst)), C.c_uint32(flags), ok, nthreads)
    return ok.raw


def verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, group=256, rand_bits=64, flags=0, nthreads=1):
    """one accept byte per group of `group` consecutive tuples (include/bls_bn254_b200.h: bn254_verify_batch_randomized)"""
    blob, off = pack_msgs(msgs)
    n = len(msgs)
    ng = (n + group - 1) // group
    ok = _buf(max(ng, 1))
    rc = lib().orc_verify_batch_randomized(variant, pks, blob, off, sigs, C.c_size_t(n), dst, C.c_size_t(len(dst)), C.c_uint32(flags), seed,
                                           rand_bits, C.c_size_t(group), ok, nthreads)
    assert rc == 0, rc
    return ok.raw[:ng]


def aggregate_verify(variant, pks, msgs, agg_sig, dst, flags=0, nthreads=1):
    blob, off = pack_msgs(msgs)
    ok = _buf(1)
    lib().orc_aggregate_verify(variant, pks, blob, off, C.c_size_t(len(msgs)), agg_sig, dst, C.c_size_t(len(dst)), C.c_uint32(flags), ok, nthreads)
    return ok.raw[0]


def g1_mul(pts, scalars, nthreads=1):
    n = len(pts) // 64
    out = _buf(64 * n)
    lib().orc_g1_mul(pts, scalars, C.c_size_t(n), out, nthreads)
    return out.raw


def g2_mul(pts, scalars, nthreads=1):
    n = len(pts) // 128
    out = _buf(128 * n)
    lib().orc_g2_mul(pts, scalars, C.c_size_t(n), out, nthreads)
    return out.raw


def g1_sum(pts):
    out = _buf(64)
    assert lib().orc_g1_sum(pts, C.c_size_t(len(pts) // 64), out) == 0
    return out.raw


def fast_aggregate_verify_batch(variant, pks, key_counts, msgs, sigs, dst, flags=0):
    """one accept byte per group: group g = the next key_counts[g] public keys, message g, aggregate signature g"""
    blob, off = pack_msgs(msgs)
    koff = (C.c_uint64 * (len(key_counts) + 1))()
    for i, k in enumerate(key_counts):
        koff[i + 1] = koff[i] + k
    ok = _buf(max(1, len(msgs)))
    lib().orc_fast_aggregate_verify_batch(variant, pks, koff, blob, off, sigs, C.c_size_t(len(msgs)), dst, C.c_size_t(len(dst)), C.c_uint32(flags), ok)
    return ok.raw[:len(msgs)]


def verify_same_key_randomized(variant, pks, msg_counts, msgs, sigs, dst, seed, rand_bits=64, flags=0, nthreads=1):
    """one accept byte per KEY: key k owns the next msg_counts[k] (message, signature) tuples"""
    blob, off = pack_msgs(msgs)
    nk = len(msg_counts)
    koff = (C.c_uint64 * (nk + 1))()
    for i, k in enumerate(msg_counts):
        koff[i + 1] = koff[i] + k
    assert koff[nk] == len(msgs)
    ok = _buf(max(1, nk))
    rc = lib().orc_verify_same_key_randomized(variant, pks, koff, C.c_size_t(nk), blob, off, sigs, dst, C.c_size_t(len(dst)), C.c_uint32(flags),
                                              seed, rand_bits, ok, nthreads)
    assert rc == 0, rc
    return ok.raw[:nk]


def g2_sum(pts):
    out = _buf(128)
    assert lib().orc_g2_sum(pts, C.c_size_t(len(pts) // 128), out) == 0
    return out.raw


def g1_generator():
    out = _buf(64)
    lib().orc_g1_generator(out)
    return out.raw


def g2_generator():
    out = _buf(128)
    lib().orc_g2_generator(out)
    return out.raw


def g1_compress(pts):
    n = len(pts) // 64
    out = _buf(32 * n)
    lib().orc_g1_compress(pts, C.c_size_t(n), out)
    return out.raw


def g1_decompress(comp, as_written=0):
    n = len(comp) // 32
    out, st = _buf(64 * n), _buf(n)
    lib().orc_g1_decompress(comp, C.c_size_t(n), out, st, as_written)
    return out.raw, st.raw


def g2_compress(pts):
    n = len(pts) // 128
    out = _buf(64 * n)
    lib().orc_g2_compress(pts, C.c_size_t(n), out)
    return out.raw


def g2_decompress(comp):
    n = len(comp) // 64
    out, st = _buf(128 * n), _buf(n)
    lib().orc_g2_decompress(comp, C.c_size_t(n), out, st)
    return out.raw, st.raw


def g1_is_on_curve(pts):
    n = len(pts) // 64
    out = _buf(n)
    lib().orc_g1_is_on_curve(pts, C.c_size_t(n), out)
    return out.raw


def g2_is_on_curve(pts):
    n = len(pts) // 128
    out = _buf(n)
    lib().orc_g2_is_on_curve(pts, C.c_size_t(n), out)
    return out.raw


def g1_is_torsion_free(pts):
    n = len(pts) // 64
    out = _buf(n)
    lib().orc_g1_is_torsion_free(pts, C.c_size_t(n), out)
    return out.raw


def g2_is_torsion_free(pts):
    n = len(pts) // 128
    out = _buf(n)
    lib().orc_g2_is_torsion_free(pts, C.c_size_t(n), out)
    return out.raw


def mul_count(reset=False):
    v = lib().orc_mul_count()
    if reset:
        lib().orc_mul_count_reset()
    return v
