"""CPU suite for the DEVICE code: tests/hostsim compiles bls_bn254_b200/csrc/*.cuh with g++ (carry
chains emulated in C) and the very per-item bodies the CUDA kernels run are compared with the oracle
and the reference's golden vectors.  Test infrastructure only: the product never links this build."""
import ctypes as C
import pathlib
import random
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
QUUX_G1 = b"QUUX-V01-CS02-with-BN254G1_XMD:SHA-256_SVDW_RO_"
QUUX_G2 = b"QUUX-V01-CS02-with-BN254G2_XMD:SHA-256_SVDW_RO_"
H = bytes.fromhex


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "tests" / "hostsim")])
    return C.CDLL(str(ROOT / "tests" / "hostsim" / "libhostsim.so"))


def _w(x):
    return (C.c_uint32 * 8)(*[(x >> (32 * i)) & 0xffffffff for i in range(8)])


def _r(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


def test_fp_arithmetic(sim):
    rng = random.Random(1)
    ri = pow(1 << 256, -1, P)
    vals = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 1 << 253, (1 << 253) - 1, 0xffffffff, 0xffffffff << 224 & (P - 1)] + [rng.randrange(P) for _ in range(400)]
    for _ in range(4000):
        a, b = rng.choice(vals), rng.choice(vals)
        o = (C.c_uint32 * 8)()
        sim.sim_fp_mul(o, _w(a), _w(b))
        assert _r(o) == a * b * ri % P
        sim.sim_fp_add(o, _w(a), _w(b))
        assert _r(o) == (a + b) % P
        sim.sim_fp_sub(o, _w(a), _w(b))
        assert _r(o) == (a - b) % P


def test_fp2_sqrt(sim):
    from oracle import pymodel as pm
    rng = random.Random(2)
    cases = [(0, 0), (4, 0), (P - 4, 0), (0, 9), (3, 0), (P - 3, 0)]
    for _ in range(60):
        x = (rng.randrange(P), rng.randrange(P))
        cases.append(pm.f2_sqr(x))
        cases.append(x)
    for a in cases:
        out = C.create_string_buffer(64)
        ok = sim.sim_fp2_sqrt(out, a[0].to_bytes(32, "big") + a[1].to_bytes(32, "big"))
        assert bool(ok) == pm.f2_is_square(a) or a == (0, 0)
        if ok:
            r = (int.from_bytes(out.raw[:32], "big"), int.from_bytes(out.raw[32:], "big"))
            assert pm.f2_sqr(r) == a


def _hash(sim, fn, rec, msgs, dst, count, orc):
    blob, off = orc.pack_msgs(msgs)
    out, st = C.create_string_buffer(rec * len(msgs)), C.create_string_buffer(len(msgs))
    fn(blob, off, C.c_size_t(len(msgs)), dst, C.c_size_t(len(dst)), count, out, st)
    return out.raw, st.raw


def test_hash_kats_and_oracle(sim, orc, kats):
    for suite, count in (("g1_encode", 1), ("g1_hash", 2)):
        dst = kats[suite + "_dst"].encode()
        msgs = [v["msg"].encode("latin1") for v in kats[suite]]
        out, st = _hash(sim, sim.sim_hash_to_g1, 64, msgs, dst, count, orc)
        for i, v in enumerate(kats[suite]):
            assert out[64 * i:64 * i + 64] == H(v["p_x"]) + H(v["p_y"])
    for suite, count in (("g2_encode", 1), ("g2_hash", 2)):
        dst = kats[suite + "_dst"].encode()
        msgs = [v["msg"].encode("latin1") for v in kats[suite]]
        out, st = _hash(sim, sim.sim_hash_to_g2, 128, msgs, dst, count, orc)
        for i, v in enumerate(kats[suite]):
            assert out[128 * i:128 * i + 128] == H(v["x_c1"]) + H(v["x_c0"]) + H(v["y_c1"]) + H(v["y_c0"])
    rng = random.Random(3)
    msgs = [rng.randbytes(rng.randrange(0, 150)) for _ in range(40)]
    for dst in (QUUX_G1, b"", b"d" * 255):
        assert _hash(sim, sim.sim_hash_to_g1, 64, msgs, dst, 2, orc)[0] == orc.hash_to_g1(msgs, dst)
    assert _hash(sim, sim.sim_hash_to_g2, 128, msgs[:12], QUUX_G2, 2, orc)[0] == orc.hash_to_g2(msgs[:12], QUUX_G2)
    for dst in (b"short", b"L" * 256, b"L" * 1000):
        out = C.create_string_buffer(96)
        sim.sim_expand_message_xmd(b"msg", C.c_size_t(3), dst, C.c_size_t(len(dst)), out, C.c_size_t(96))
        assert out.raw == orc.expand_message_xmd(b"msg", dst, 96)


def test_pairing_vs_oracle(sim, orc, kats):
    rng = random.Random(4)
    msgs = [rng.randbytes(8) for _ in range(4)]
    g1 = orc.g1_generator() + orc.hash_to_g1(msgs, QUUX_G1)
    g2 = orc.g2_generator() + orc.hash_to_g2(msgs, QUUX_G2)
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    g1 += ident1 + b"\xff" * 64
    g2 += g2[:128] * 2
    n = len(g1) // 64
    for mode in (1, 0):
        out, st = C.create_string_buffer(384 * n), C.create_string_buffer(n)
        sim.sim_pairing_batch(g1, g2, C.c_size_t(n), out, st, mode)
        ref, rst = (orc.pairing_batch(g1, g2, want_status=True) if mode == 0 else (orc.miller_loop_batch(g1, g2), None))
        assert out.raw == ref
        if rst is not None:
            assert st.raw == rst
        if mode == 0:
            assert out.raw[:384] == b"".join(H(x) for x in kats["gt_generator"])  # pairings.rs:970-980
    ml = orc.miller_loop_batch(g1[:192], g2[:384])
    out = C.create_string_buffer(len(ml))
    sim.sim_final_exp(ml, C.c_size_t(3), out)
    assert out.raw == orc.final_exponentiation(ml)


def test_two_pair_miller_loops(sim, orc):
    """shared-squaring 2-pair loops: variable/variable and (P0, -g2) through the precomputed line table"""
    rng = random.Random(5)
    msgs = [rng.randbytes(8) for _ in range(2)]
    g1 = orc.hash_to_g1(msgs, QUUX_G1)
    g2 = orc.hash_to_g2(msgs, QUUX_G2)
    out = C.create_string_buffer(384)
    sim.sim_miller2(g1[:64], g2[:128], g1[64:], g2[128:], 0, out)
    assert out.raw == orc.multi_miller_loop(g1, g2)
    gen = orc.g2_generator()
    neg = gen[:64] + ((P - int.from_bytes(gen[64:96], "big")) % P).to_bytes(32, "big") + ((P - int.from_bytes(gen[96:], "big")) % P).to_bytes(32, "big")
    sim.sim_miller2(g1[:64], None, g1[64:], g2[128:], 1, out)
    assert out.raw == orc.multi_miller_loop(g1, neg + g2[128:])
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    sim.sim_miller2(ident1, None, g1[64:], g2[128:], 1, out)
    assert out.raw == orc.multi_miller_loop(ident1 + g1[64:], neg + g2[128:]) == orc.miller_loop_batch(g1[64:], g2[128:])


def test_scalar_mul_and_codecs(sim, orc):
    rng = random.Random(6)
    msgs = [rng.randbytes(8) for _ in range(6)]
    g1 = orc.g1_generator() + orc.hash_to_g1(msgs, QUUX_G1)
    g2 = orc.g2_generator() + orc.hash_to_g2(msgs, QUUX_G2)
    n = 7
    sc = (0).to_bytes(32, "big") + (1).to_bytes(32, "big") + b"".join(rng.randrange(1 << 255).to_bytes(32, "big") for _ in range(n - 2))
    o1, o2 = C.create_string_buffer(64 * n), C.create_string_buffer(128 * n)
    sim.sim_g1_mul(g1, sc, C.c_size_t(n), o1)
    sim.sim_g2_mul(g2, sc, C.c_size_t(n), o2)
    assert o1.raw == orc.g1_mul(g1, sc)
    assert o2.raw == orc.g2_mul(g2, sc)
    comp, back, st = C.create_string_buffer(32 * n), C.create_string_buffer(64 * n), C.create_string_buffer(n)
    sim.sim_g1_codec(g1, C.c_size_t(n), comp, back, st)
    assert comp.raw == orc.g1_compress(g1) and back.raw == g1 and st.raw == b"\x01" * n
    comp, back = C.create_string_buffer(64 * n), C.create_string_buffer(128 * n)
    sim.sim_g2_codec(g2, C.c_size_t(n), comp, back, st)
    assert comp.raw == orc.g2_compress(g2) and back.raw == g2 and st.raw == b"\x01" * n


def test_fp12_tower_operations_one_by_one(sim, orc):
    """Fp6 / Fp12 product, square, inverse, conjugate, the three Frobenius maps, the Granger-Scott cyclotomic square and the
    sparse 0/3/4 line product (single and paired) of the device code against the big-integer model (oracle/pymodel.py) -- the operations no
    reference test pins individually -- plus the identities x * x^-1 = 1 and frob_k(x) = x^(p^k)"""
    from oracle import pymodel as pm
    rng = random.Random(31)

    def rnd12():
        return tuple(tuple((rng.randrange(P), rng.randrange(P)) for _ in range(3)) for _ in range(2))

    def call(op, a, b=None):
        """one thread per element, and the lane-pair representation (two emulated lanes): both must give the same bytes"""
        out, out2 = C.create_string_buffer(384), C.create_string_buffer(384)
        assert sim.sim_fp12_op(op, pm.gt_to_repr(a), pm.gt_to_repr(b) if b is not None else None, out) == 0
        assert sim.sim_fp12_op_l2(op, pm.gt_to_repr(a), pm.gt_to_repr(b) if b is not None else None, out2) == 0
        assert out.raw == out2.raw, "lane-pair result differs (op %d)" % op
        return out.raw

    one = pm.gt_to_repr(pm.F12_ONE)
    for _ in range(4):
        x, y = rnd12(), rnd12()
        assert call(0, x, y) == pm.gt_to_repr(pm.f12_mul(x, y))
        assert call(1, x) == pm.gt_to_repr(pm.f12_sqr(x))
        inv = pm.f12_inv(x)
        assert call(2, x) == pm.gt_to_repr(inv)
        assert call(0, x, inv) == one
        assert call(3, x) == pm.gt_to_repr(pm.f12_conj(x))
        for k in (1, 2, 3):
            assert call(3 + k, x) == pm.gt_to_repr(pm.f12_frob(x, k))
        c0, c3, c4 = y[0]
        assert call(8, x, y) == pm.gt_to_repr(pm.f12_mul(x, pm.f12_sparse034(c0, c3, c4)))
        # the paired product of the two-pair Miller loops: x * l * l' (sparse by sparse first), also with a neutral / a zero line
        d0, d3, d4 = y[1]
        assert call(9, x, y) == pm.gt_to_repr(pm.f12_mul(pm.f12_mul(x, pm.f12_sparse034(c0, c3, c4)), pm.f12_sparse034(d0, d3, d4)))
        neutral = (y[0], ((1, 0), (0, 0), (0, 0)))
        assert call(9, x, neutral) == call(8, x, y)
        assert call(9, x, (y[0], ((0, 0), (0, 0), (0, 0)))) == pm.gt_to_repr(tuple(tuple((0, 0) for _ in range(3)) for _ in range(2)))
    x = rnd12()
    assert call(4, x) == pm.gt_to_repr(pm.f12_pow(x, P))          # Frobenius really is the p-th power
    # cyclotomic subgroup: x^((p^6 - 1)(p^2 + 1)); there squaring has the compressed Granger-Scott form and conj = inverse
    e = pm.f12_mul(pm.f12_conj(x), pm.f12_inv(x))
    e = pm.f12_mul(pm.f12_frob(e, 2), e)
    assert call(7, e) == pm.gt_to_repr(pm.f12_sqr(e)) == call(1, e)
    assert call(0, e, pm.f12_conj(e)) == one
    # edge values
    zero12 = tuple(tuple((0, 0) for _ in range(3)) for _ in range(2))
    assert call(0, x, zero12) == pm.gt_to_repr(zero12) and call(0, x, pm.F12_ONE) == pm.gt_to_repr(x)
    assert call(2, zero12) == pm.gt_to_repr(zero12)               # 0 -> 0 like the field inversion underneath


def test_subgroup_checks_without_the_big_multiplication(sim, orc):
    """is_torsion_free as the kernels evaluate it (G1: cofactor 1, so on-curve; G2: [x+1]Q + psi([x]Q) + psi^2([x]Q) ==
    psi^3([2x]Q)) against the oracle's [r]P == O (g1.rs:852-855, g2.rs:733-736) on subgroup points, points of the twist
    outside the subgroup, points of the cofactor subgroup, sums of both, the identity, off-curve and undecodable bytes"""
    from oracle import pymodel as pm
    rng = random.Random(21)
    msgs = [rng.randbytes(7) for _ in range(5)]
    g1 = orc.g1_generator() + orc.hash_to_g1(msgs, QUUX_G1)
    off1 = bytearray(g1[:64]); off1[40] ^= 1
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    pts1 = g1 + bytes(off1) + ident1 + b"\xff" * 64
    good2 = orc.g2_generator() + orc.hash_to_g2(msgs, QUUX_G2)
    raw = orc.svdw_g2(b"".join((i + 5).to_bytes(32, "big") + (i + 9).to_bytes(32, "big") for i in range(4)))   # on the twist, not in G2
    cof = orc.g2_mul(raw, pm.R_ORDER.to_bytes(32, "big") * 4)                                                   # cofactor-subgroup components
    mixed = b"".join(orc.g2_sum(good2[128 * i:128 * (i + 1)] + cof[128 * i:128 * (i + 1)]) for i in range(4))
    off2 = bytearray(good2[:128]); off2[100] ^= 1
    ident2 = bytes(64) + bytes(63) + b"\x01"
    pts2 = good2 + raw + cof + mixed + bytes(off2) + ident2 + b"\xff" * 128
    n1, n2 = len(pts1) // 64, len(pts2) // 128
    o1, o2 = C.create_string_buffer(n1), C.create_string_buffer(n2)
    sim.sim_torsion_free(pts1, C.c_size_t(n1), pts2, C.c_size_t(n2), o1, o2)
    assert o1.raw == orc.g1_is_torsion_free(pts1) == b"\x01" * 6 + b"\x00\x01\x00"
    assert o2.raw == orc.g2_is_torsion_free(pts2) == b"\x01" * 6 + b"\x00" * 12 + b"\x00\x01\x00"


def test_is_square_jacobi(sim):
    """fp_is_square as a binary Jacobi symbol == the Euler criterion it replaces == Python's pow, on edge values
    (0, 1, 2, p-1, powers of two, values with long runs of zero bits) and 4000 random residues"""
    rng = random.Random(12)
    vals = [0, 1, 2, 3, 4, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 32, 1 << 64, (1 << 64) + 1, 1 << 253, (1 << 253) + (1 << 32),
            (1 << 200) - (1 << 100), 3 << 96, 0xffffffff, 0x100000000 * 5]
    vals += [rng.randrange(P) for _ in range(4000)]
    vals += [rng.randrange(1 << rng.randrange(1, 254)) for _ in range(500)]
    vals += [(rng.randrange(1 << 60) << rng.randrange(1, 190)) % P for _ in range(500)]
    n = len(vals)
    buf = (C.c_uint32 * (8 * n))(*[(v >> (32 * j)) & 0xffffffff for v in vals for j in range(8)])
    jac, eul = C.create_string_buffer(n), C.create_string_buffer(n)
    sim.sim_fp_is_square(buf, C.c_size_t(n), jac, eul)
    want = bytes(1 if v == 0 or pow(v, (P - 1) // 2, P) == 1 else 0 for v in vals)
    assert jac.raw == want
    assert eul.raw == want
    assert 0.4 < sum(want) / n < 0.6


def test_inversion_binary_gcd(sim):
    """fp_inv as a binary extended GCD == the Fermat exponentiation it replaces == Python's pow(x, -1, p), on edge values
    (0 -> 0, 1, p-1, powers of two, tiny and huge values: the shift count k covers both post-processing branches)"""
    rng = random.Random(13)
    R = 1 << 256
    vals = [0, 1, 2, 3, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 32, 1 << 64, 1 << 253, (1 << 253) + 1, 3 << 96, 0xffffffff, R % P, R * R % P]
    vals += [rng.randrange(1, P) for _ in range(3000)]
    vals += [rng.randrange(1, 1 << rng.randrange(1, 254)) for _ in range(500)]
    vals += [(rng.randrange(1, 1 << 60) << rng.randrange(1, 190)) % P for _ in range(300)]
    vals += [P - rng.randrange(1, 1 << rng.randrange(1, 200)) for _ in range(300)]
    n = len(vals)
    buf = (C.c_uint32 * (8 * n))(*[(v >> (32 * j)) & 0xffffffff for v in vals for j in range(8)])
    o1, o2 = (C.c_uint32 * (8 * n))(), (C.c_uint32 * (8 * n))()
    sim.sim_fp_inv(buf, C.c_size_t(n), o1, o2)
    for i, v in enumerate(vals):
        got1 = sum(int(o1[8 * i + j]) << (32 * j) for j in range(8))
        got2 = sum(int(o2[8 * i + j]) << (32 * j) for j in range(8))
        want = pow(v, -1, P) * R * R % P if v else 0      # Montgomery form in, Montgomery form out
        assert got1 == want, hex(v)
        assert got2 == want, hex(v)


def test_randomized_batch_scalars_and_scalings(sim, orc):
    """rbv_scalar = the two halves of SHA-256(seed || LE64(i)) (hashlib); the two scalings of a tuple (affine [r]P for
    the pairing, projective [r]S summed over the group), done on the device by the GLV joint double-and-add, against
    the oracle's PLAIN scalar multiplication by the integer r = a + b lambda mod r"""
    import hashlib
    from oracle import pymodel as pm
    rng = random.Random(8)
    seed = rng.randbytes(32)
    kw = (C.c_uint32 * 4)()
    for idx in (0, 1, 255, 256, 2**32 + 5, 2**63 + 11):
        dg = hashlib.sha256(seed + idx.to_bytes(8, "little")).digest()
        for bits in (64, 128):
            h = bits // 16
            sim.sim_rbv_scalar(seed, C.c_uint64(idx), bits, kw)
            assert int(kw[0]) | int(kw[1]) << 32 == int.from_bytes(dg[:h], "big")
            assert int(kw[2]) | int(kw[3]) << 32 == int.from_bytes(dg[h:2 * h], "big")
    n = 5
    msgs = [rng.randbytes(9) for _ in range(n - 1)]
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    ident2 = bytes(64) + bytes(63) + b"\x01"
    g1 = orc.hash_to_g1(msgs, QUUX_G1) + ident1
    g2 = orc.hash_to_g2(msgs[:-1], QUUX_G2) + ident2 + orc.g2_generator()
    for bits, idx0 in ((64, 0), (128, 1000)):
        rs = b"".join(pm.rbv_scalar(seed, idx0 + i, bits).to_bytes(32, "big") for i in range(n))
        o1, s1, s2 = C.create_string_buffer(64 * n), C.create_string_buffer(64), C.create_string_buffer(128)
        sim.sim_rbv_scale(g1, g2, C.c_size_t(n), seed, C.c_uint64(idx0), bits, o1, s1, s2)
        assert o1.raw == orc.g1_mul(g1, rs)
        assert s1.raw == orc.g1_sum(orc.g1_mul(g1, rs))
        assert s2.raw == orc.g2_sum(orc.g2_mul(g2, rs))


def test_lane_pair_pairing(sim, orc):
    """two lanes per pairing (lane-split Fp2, BN_LANES=2 kernels): the same bodies run by two emulated
    lanes exchanging through the shuffle hook must equal the oracle, identity and bad inputs included"""
    rng = random.Random(4)
    msgs = [rng.randbytes(8) for _ in range(3)]
    g1 = orc.g1_generator() + orc.hash_to_g1(msgs, QUUX_G1)
    g2 = orc.g2_generator() + orc.hash_to_g2(msgs, QUUX_G2)
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    g1 += ident1 + b"\xff" * 64
    g2 += g2[:256]
    n = len(g1) // 64
    for mode in (1, 0):
        out, st = C.create_string_buffer(384 * n), C.create_string_buffer(n)
        sim.sim_pairing_batch_l2(g1, g2, C.c_size_t(n), out, st, mode)
        assert out.raw == (orc.pairing_batch(g1, g2) if mode == 0 else orc.miller_loop_batch(g1, g2))
    out = C.create_string_buffer(384)
    sim.sim_miller2_l2(g1[64:128], g2[128:256], g1[128:192], g2[256:384], 0, out)
    assert out.raw == orc.multi_miller_loop(g1[64:192], g2[128:384])
    gen = orc.g2_generator()
    neg = gen[:64] + ((P - int.from_bytes(gen[64:96], "big")) % P).to_bytes(32, "big") + ((P - int.from_bytes(gen[96:], "big")) % P).to_bytes(32, "big")
    sim.sim_miller2_l2(g1[64:128], None, g1[128:192], g2[256:384], 1, out)
    assert out.raw == orc.multi_miller_loop(g1[64:192], neg + g2[256:384])
    ml = orc.miller_loop_batch(g1[:128], g2[:256])
    o2 = C.create_string_buffer(len(ml))
    sim.sim_final_exp_l2(ml, C.c_size_t(2), o2)
    assert o2.raw == orc.final_exponentiation(ml)


def test_fused_products(sim):
    """fused two-term dot product and the Fp2 product built from it (raw limbs)"""
    rng = random.Random(7)
    R = 1 << 256
    ri = pow(R, -1, P)

    def w(x, n=8):
        return (C.c_uint32 * n)(*[(x >> (32 * i)) & 0xffffffff for i in range(n)])

    def r(a, n=8):
        return sum(int(a[i]) << (32 * i) for i in range(n))
    fv = [0, 1, P - 1, P - 2, (P - 1) // 2, 2**253] + [rng.randrange(P) for _ in range(60)]
    for _ in range(2000):
        a, b, c, d = [rng.choice(fv) for _ in range(4)]
        o = (C.c_uint32 * 8)()
        sim.sim_fp_dot2(o, w(a), w(b), w(c), w(d))
        assert r(o) == (a * b + c * d) * ri % P
        o = (C.c_uint32 * 16)()
        sim.sim_fp2_mul(o, w(a | (b << 256), 16), w(c | (d << 256), 16))
        assert r(o, 16) == ((a * c - b * d) * ri % P) | (((a * d + b * c) * ri % P) << 256)


def test_prepared_keys(sim, orc):
    """G2Prepared (pairings.rs:609-660): the line tables equal the oracle's byte for byte, and the two-pair loop fed from the
    table equals multi_miller_loop in both shapes, a dead pair included"""
    rng = random.Random(3)
    msgs = [rng.randbytes(9) for _ in range(3)]
    g1, g2 = orc.hash_to_g1(msgs, QUUX_G1), orc.hash_to_g2(msgs, QUUX_G2)
    out = C.create_string_buffer(88 * 192)
    for i in range(3):
        sim.sim_g2_prepare(g2[128 * i:128 * (i + 1)], out)
        assert out.raw == orc.g2_prepare(g2[128 * i:128 * (i + 1)])[0]
    g = orc.g2_generator()
    neg = lambda b: ((P - int.from_bytes(b, "big")) % P).to_bytes(32, "big")
    ng2 = g[:64] + neg(g[64:96]) + neg(g[96:])
    ident = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    o = C.create_string_buffer(384)
    for lanes in (0, 1):
        sim.sim_miller2_prepared(g1[:64], g1[64:128], g2[128:256], lanes, o)
        assert o.raw == orc.multi_miller_loop(g1[:128], ng2 + g2[128:256])
        sim.sim_miller2_prepared(g1[:64], ident, g2[128:256], lanes, o)
        assert o.raw == orc.multi_miller_loop(g1[:64] + ident, ng2 + g2[128:256])
