"""Pure-Python big-integer model of the BN254 hot path (TEST INFRASTRUCTURE ONLY).

This file is part of the oracle: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it.  It is the slow, obviously-correct third leg next to the C
oracle (oracle/bn254_oracle.c) and the CUDA product; it also derives every constant the
other two use (tools/gen_constants.py).  Small cases only (one pairing ~ 0.3 s).

Follows, by file:line under /root/reference/src/inner_types:
  fp.rs:115-122,284-371,428-458   from_okm, SvdW map to G1, is_square, hash/encode
  fp2.rs:95-99,161-287,377-488    sgn0, invert, sqrt, SvdW map to G2, arithmetic, hash/encode
  g1.rs:104-118,283-360,910-928   affine conversion, codecs, hash/encode
  g2.rs:98-112,274-388,685-693,919-954  affine conversion, codecs, clear_cofactor, psi, hash
  pairings.rs:499-514,760-962     Gt::to_repr, pairing, doubling_step/addition_step/ell
Pairing convention ("Tier B"): the function the reference's own
known-answer test demands (pairings.rs:387-479,970-980): optimal ate over the tower
Fp2=Fp[u]/(u^2+1), Fp6=Fp2[v]/(v^3-(9+u)), Fp12=Fp6[w]/(w^2-v), Miller loop over NAF(6x+2)
with the reference's own doubling_step/addition_step line formulas placed at w^0,w^1,w^3,
final exponent (p^12-1)/r * 2x(6x^2+3x+1).
"""
import hashlib

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R_ORDER = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
X = 0x44e992b44a6909f1
ATE = 6 * X + 2
assert P == 36 * X**4 + 36 * X**3 + 24 * X**2 + 6 * X + 1
assert R_ORDER == 36 * X**4 + 36 * X**3 + 18 * X**2 + 6 * X + 1

# ---------------------------------------------------------------- Fp2
F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (9, 1)


def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_conj(a): return (a[0], (-a[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_muls(a, s): return (a[0] * s % P, a[1] * s % P)
def f2_mul_xi(a): return ((9 * a[0] - a[1]) % P, (a[0] + 9 * a[1]) % P)


def f2_inv(a):
    t = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * t % P, (-a[1]) * t % P)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def f2_sgn0(a):  # fp2.rs:95-99
    return (a[0] & 1) | ((a[0] == 0) & (a[1] & 1))


def f2_is_square(a):  # fp2.rs:441-452
    t = pow((a[0] * a[0] + a[1] * a[1]) % P, (P - 1) // 2, P)
    return t in (0, 1)


def f2_sqrt(a):  # fp2.rs:172-218 (Alg. 9 of eprint 2012/685); returns None if not a square
    if a == F2_ZERO:
        return F2_ZERO
    a1 = f2_pow(a, (P - 3) // 4)
    alpha = f2_mul(f2_sqr(a1), a)
    x0 = f2_mul(a1, a)
    if alpha == ((-1) % P, 0):
        s = ((-x0[1]) % P, x0[0])
    else:
        s = f2_mul(f2_pow(f2_add(alpha, F2_ONE), (P - 1) // 2), x0)
    return s if f2_sqr(s) == a else None


# ---------------------------------------------------------------- Fp6 = Fp2[v]/(v^3 - xi)
F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a):  # multiply by v
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


# ---------------------------------------------------------------- Fp12 = Fp6[w]/(w^2 - v)
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return (f6_add(f6_mul(a0, b0), f6_mul_v(f6_mul(a1, b1))),
            f6_add(f6_mul(a0, b1), f6_mul(a1, b0)))


def f12_sqr(a): return f12_mul(a, a)
def f12_conj(a): return (a[0], f6_neg(a[1]))


def f12_inv(a):
    a0, a1 = a
    t = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))))
    return (f6_mul(a0, t), f6_neg(f6_mul(a1, t)))


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


def frob_gamma(k, i):
    """xi^(i*(p^k-1)/6): coefficient of w^i under the p^k-power Frobenius."""
    return f2_pow(XI, i * (P**k - 1) // 6)


_GAMMA = {k: [frob_gamma(k, i) for i in range(6)] for k in (1, 2, 3)}


def f12_frob(a, k=1):
    (a0, a2, a4), (a1, a3, a5) = a
    co = [a0, a1, a2, a3, a4, a5]
    if k & 1:
        co = [f2_conj(c) for c in co]
    co = [f2_mul(c, _GAMMA[k][i]) for i, c in enumerate(co)]
    return ((co[0], co[2], co[4]), (co[1], co[3], co[5]))


def f12_sparse034(c0, c3, c4):
    """Fp12 element with c0.c0=c0, c1.c0=c3, c1.c1=c4 (w^0, w^1, w^3), everything else 0."""
    return ((c0, F2_ZERO, F2_ZERO), (c3, c4, F2_ZERO))


def gt_to_repr(f):  # pairings.rs:499-514
    out = b""
    for c6 in f:
        for c2 in c6:
            out += c2[0].to_bytes(32, "big") + c2[1].to_bytes(32, "big")
    return out


# ---------------------------------------------------------------- curves (affine, None = identity)
B1 = 3
B2 = f2_mul((3, 0), f2_inv(XI))
G1_GEN = (1, 2)
G2_GEN = ((0x1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed,
           0x198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2),
          (0x12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa,
           0x090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b))


class _Fp:
    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    mul = staticmethod(lambda a, b: a * b % P)
    inv = staticmethod(lambda a: pow(a, -1, P))
    neg = staticmethod(lambda a: (-a) % P)
    zero = 0


class _Fp2:
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    mul = staticmethod(f2_mul)
    inv = staticmethod(f2_inv)
    neg = staticmethod(f2_neg)
    zero = F2_ZERO


def ec_add(F, a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        num = F.mul(F.add(F.add(x1, x1), x1), x1)
        lam = F.mul(num, F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    return (x3, F.sub(F.mul(lam, F.sub(x1, x3)), y1))


def ec_neg(F, a):
    return None if a is None else (a[0], F.neg(a[1]))


def ec_mul(F, a, k):
    if k < 0:
        return ec_mul(F, ec_neg(F, a), -k)
    r = None
    for bit in bin(k)[2:] if k else "":
        r = ec_add(F, r, r)
        if bit == "1":
            r = ec_add(F, r, a)
    return r


def g1_add(a, b): return ec_add(_Fp, a, b)
def g1_mul(a, k): return ec_mul(_Fp, a, k)
def g1_neg(a): return ec_neg(_Fp, a)
def g2_add(a, b): return ec_add(_Fp2, a, b)
def g2_mul(a, k): return ec_mul(_Fp2, a, k)
def g2_neg(a): return ec_neg(_Fp2, a)
def g1_on_curve(a): return a is None or (a[1] * a[1] - a[0] ** 3 - B1) % P == 0
def g2_on_curve(a): return a is None or f2_sqr(a[1]) == f2_add(f2_mul(f2_sqr(a[0]), a[0]), B2)


# psi constants, g2.rs:939-947 (= xi^((p-1)/3), xi^((p-1)/2))
ENDO_U = f2_pow(XI, (P - 1) // 3)
ENDO_V = f2_pow(XI, (P - 1) // 2)


def g2_psi(a):  # g2.rs:938-954
    if a is None:
        return None
    return (f2_mul(f2_conj(a[0]), ENDO_U), f2_mul(f2_conj(a[1]), ENDO_V))


def g2_clear_cofactor(a):  # g2.rs:685-693
    p0 = g2_mul(a, X)
    p1 = g2_psi(g2_add(g2_add(p0, p0), p0))
    p2 = g2_psi(g2_psi(p0))
    p3 = g2_psi(g2_psi(g2_psi(a)))
    return g2_add(g2_add(g2_add(p0, p1), p2), p3)


# ---------------------------------------------------------------- hash to field / curve
def expand_message_xmd(msg: bytes, dst: bytes, n: int) -> bytes:
    """RFC 9380 section 5.3.1 with SHA-256 (elliptic-curve 0.13.8 ExpandMsgXmd, external to the
    reference tree; call sites fp.rs:433-458, fp2.rs:463-488)."""
    if len(dst) > 255:
        dst = hashlib.sha256(b"H2C-OVERSIZE-DST-" + dst).digest()
    ell = (n + 31) // 32
    assert ell <= 255 and n <= 65535
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(b"\x00" * 64 + msg + n.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:n]


def fp_from_okm(b48: bytes) -> int:  # fp.rs:115-122
    return int.from_bytes(b48, "big") % P


def hash_to_field_fp(msg, dst, count):  # fp.rs:433-458
    okm = expand_message_xmd(msg, dst, 48 * count)
    return [fp_from_okm(okm[48 * i:48 * i + 48]) for i in range(count)]


def hash_to_field_fp2(msg, dst, count):  # fp2.rs:454-488
    okm = expand_message_xmd(msg, dst, 96 * count)
    return [(fp_from_okm(okm[96 * i:96 * i + 48]), fp_from_okm(okm[96 * i + 48:96 * i + 96]))
            for i in range(count)]


SVDW1_C1 = 4
SVDW1_C2 = (P - 1) // 2            # -Z/2 = -1/2
SVDW1_C3 = 0x016789af3a83522eb353c98fc6b36d713d5d8d1cc5dffffffa
SVDW1_C4 = 0x10216f7ba065e00de81ac1e7808072c9dd2b2385cd7b438469602eb24829a9bd


def fp_is_square(a):  # fp.rs:428-431
    return pow(a, (P - 1) // 2, P) in (0, 1)


def svdw_g1(u):  # fp.rs:284-371.  Returns None where the reference panics (tv1*tv2 == 0).
    tv1 = u * u % P * SVDW1_C1 % P
    tv2 = (1 + tv1) % P
    tv1 = (1 - tv1) % P
    tv3 = tv1 * tv2 % P
    if tv3 == 0:
        return None
    tv3 = pow(tv3, -1, P)
    tv4 = u * tv1 % P * tv3 % P * SVDW1_C3 % P
    x1 = (SVDW1_C2 - tv4) % P
    gx1 = (x1 * x1 % P * x1 + 3) % P
    x2 = (SVDW1_C2 + tv4) % P
    gx2 = (x2 * x2 % P * x2 + 3) % P
    x3 = tv2 * tv2 % P * tv3 % P
    x3 = (x3 * x3 % P * SVDW1_C4 + 1) % P
    e1 = fp_is_square(gx1)
    x = x1 if e1 else x3
    if fp_is_square(gx2) and not e1:
        x = x2
    gx = (x * x % P * x + 3) % P
    y = pow(gx, (P + 1) // 4, P)
    assert y * y % P == gx
    if (u & 1) ^ (y & 1):
        y = (-y) % P
    return (x, y)


SVDW2_C1 = f2_add(F2_ONE, B2)
SVDW2_C2 = ((P - 1) // 2, 0)
SVDW2_C3 = (0x29fd332ab7260112b801fa95b21af64e2e6da55f90a3e510fcbe57377b5ca1ec,
            0x303d1eff1426764bf8408aee24ba0b865e76f77b1267a846b1e9154d01565034)
SVDW2_C4 = (0x17365bbe63b1d2078632fe0eb2ac5a41b4e6a9c08b98676721010b008d4eaf99,
            0x0f57ffe5fc79e19cd689d7aa4209cad8fe164d7f4694786b388732a995d03755)


def svdw_g2(u):  # fp2.rs:221-287
    tv1 = f2_mul(f2_sqr(u), SVDW2_C1)
    tv2 = f2_add(F2_ONE, tv1)
    tv1 = f2_sub(F2_ONE, tv1)
    tv3 = f2_mul(tv1, tv2)
    if tv3 == F2_ZERO:
        return None
    tv3 = f2_inv(tv3)
    tv4 = f2_mul(f2_mul(f2_mul(u, tv1), tv3), SVDW2_C3)
    x1 = f2_sub(SVDW2_C2, tv4)
    gx1 = f2_add(f2_mul(f2_sqr(x1), x1), B2)
    x2 = f2_add(SVDW2_C2, tv4)
    gx2 = f2_add(f2_mul(f2_sqr(x2), x2), B2)
    x3 = f2_mul(f2_sqr(tv2), tv3)
    x3 = f2_add(f2_mul(f2_sqr(x3), SVDW2_C4), F2_ONE)
    e1 = f2_is_square(gx1)
    x = x1 if e1 else x3
    if f2_is_square(gx2) and not e1:
        x = x2
    gx = f2_add(f2_mul(f2_sqr(x), x), B2)
    y = f2_sqrt(gx)
    assert y is not None
    if f2_sgn0(u) ^ f2_sgn0(y):
        y = f2_neg(y)
    return (x, y)


def hash_to_g1(msg, dst):  # g1.rs:910-919
    u0, u1 = hash_to_field_fp(msg, dst, 2)
    return g1_add(svdw_g1(u0), svdw_g1(u1))


def encode_to_g1(msg, dst):  # g1.rs:922-928
    return svdw_g1(hash_to_field_fp(msg, dst, 1)[0])


def hash_to_g2(msg, dst):  # g2.rs:919-927
    u0, u1 = hash_to_field_fp2(msg, dst, 2)
    return g2_clear_cofactor(g2_add(svdw_g2(u0), svdw_g2(u1)))


def encode_to_g2(msg, dst):  # g2.rs:930-936
    return g2_clear_cofactor(svdw_g2(hash_to_field_fp2(msg, dst, 1)[0]))


# ---------------------------------------------------------------- codecs
def g1_to_uncompressed(a):  # g1.rs:296-302 (identity = (0, 1), g1.rs:265-271)
    x, y = (0, 1) if a is None else a
    return x.to_bytes(32, "big") + y.to_bytes(32, "big")


def g1_to_compressed(a):  # g1.rs:283-288
    x, y = (0, 1) if a is None else a
    b = bytearray(x.to_bytes(32, "big"))
    b[0] |= (y & 1) << 7
    return bytes(b)


def g2_to_uncompressed(a):  # g2.rs:290-300  (c1 || c0)
    x, y = (F2_ZERO, F2_ONE) if a is None else a
    return b"".join(v.to_bytes(32, "big") for v in (x[1], x[0], y[1], y[0]))


def g2_to_compressed(a):  # g2.rs:274-283
    x, y = (F2_ZERO, F2_ONE) if a is None else a
    b = bytearray(x[1].to_bytes(32, "big") + x[0].to_bytes(32, "big"))
    b[0] |= f2_sgn0(y) << 7
    return bytes(b)


# ---------------------------------------------------------------- pairing
def naf(n):
    out = []
    while n:
        if n & 1:
            d = 2 - (n % 4)
            n -= d
        else:
            d = 0
        out.append(d)
        n >>= 1
    return out  # LSB first


ATE_NAF = naf(ATE)


def doubling_step(r):  # pairings.rs:901-930
    rx, ry, rz = r
    tmp0 = f2_sqr(rx)
    tmp1 = f2_sqr(ry)
    tmp2 = f2_sqr(tmp1)
    tmp3 = f2_sub(f2_sub(f2_sqr(f2_add(tmp1, rx)), tmp0), tmp2)
    tmp3 = f2_add(tmp3, tmp3)
    tmp4 = f2_add(f2_add(tmp0, tmp0), tmp0)
    tmp6 = f2_add(rx, tmp4)
    tmp5 = f2_sqr(tmp4)
    zsq = f2_sqr(rz)
    nx = f2_sub(f2_sub(tmp5, tmp3), tmp3)
    nz = f2_sub(f2_sub(f2_sqr(f2_add(rz, ry)), tmp1), zsq)
    ny = f2_mul(f2_sub(tmp3, nx), tmp4)
    tmp2 = f2_muls(tmp2, 8)
    ny = f2_sub(ny, tmp2)
    tmp3 = f2_mul(tmp4, zsq)
    tmp3 = f2_neg(f2_add(tmp3, tmp3))
    tmp6 = f2_sub(f2_sub(f2_sqr(tmp6), tmp0), tmp5)
    tmp6 = f2_sub(tmp6, f2_muls(tmp1, 4))
    tmp0 = f2_mul(nz, zsq)
    tmp0 = f2_add(tmp0, tmp0)
    return (nx, ny, nz), (tmp0, tmp3, tmp6)


def addition_step(r, q):  # pairings.rs:932-962
    rx, ry, rz = r
    qx, qy = q
    zsq = f2_sqr(rz)
    ysq = f2_sqr(qy)
    t0 = f2_mul(zsq, qx)
    t1 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(qy, rz)), ysq), zsq), zsq)
    t2 = f2_sub(t0, rx)
    t3 = f2_sqr(t2)
    t4 = f2_muls(t3, 4)
    t5 = f2_mul(t4, t2)
    t6 = f2_sub(f2_sub(t1, ry), ry)
    t9 = f2_mul(t6, qx)
    t7 = f2_mul(t4, rx)
    nx = f2_sub(f2_sub(f2_sub(f2_sqr(t6), t5), t7), t7)
    nz = f2_sub(f2_sub(f2_sqr(f2_add(rz, t2)), zsq), t3)
    t10 = f2_add(qy, nz)
    t8 = f2_mul(f2_sub(t7, nx), t6)
    t0 = f2_mul(ry, t5)
    t0 = f2_add(t0, t0)
    ny = f2_sub(t8, t0)
    t10 = f2_sub(f2_sqr(t10), ysq)
    t10 = f2_sub(t10, f2_sqr(nz))
    t9 = f2_sub(f2_add(t9, t9), t10)
    t10 = f2_add(nz, nz)
    t6 = f2_neg(t6)
    t1 = f2_add(t6, t6)
    return (nx, ny, nz), (t10, t1, t9)


def ell(f, coeffs, p):  # pairings.rs:888-899 with the D-twist placement 0/3/4
    c0 = f2_muls(coeffs[0], p[1])
    c1 = f2_muls(coeffs[1], p[0])
    return f12_mul(f, f12_sparse034(c0, c1, coeffs[2]))


def line_schedule(q):
    """Yield the line-coefficient triples of the Miller loop for affine Q (not identity), in the
    order the loop consumes them: per NAF digit (MSB-1 .. 0): doubling, [addition]; then the two
    Frobenius additions.  (G2Prepared, pairings.rs:609-660, done right: 65+21+2 = 88 triples.)"""
    r = (q[0], q[1], F2_ONE)
    nq = (q[0], f2_neg(q[1]))
    out = []
    for d in reversed(ATE_NAF[:-1]):
        r, c = doubling_step(r)
        out.append(c)
        if d:
            r, c = addition_step(r, q if d == 1 else nq)
            out.append(c)
    q1 = g2_psi(q)
    q2 = g2_psi(q1)
    r, c = addition_step(r, q1)
    out.append(c)
    r, c = addition_step(r, (q2[0], f2_neg(q2[1])))
    out.append(c)
    return out


def multi_miller_loop(pairs):
    """pairs: list of (G1 affine | None, G2 affine | None).  Identity terms contribute 1
    (pairings.rs:789-800, 820-835)."""
    terms = [(p, line_schedule(q)) for p, q in pairs if p is not None and q is not None]
    f = F12_ONE
    idx = 0
    for d in reversed(ATE_NAF[:-1]):
        f = f12_sqr(f)
        for p, lines in terms:
            f = ell(f, lines[idx], p)
        idx += 1
        if d:
            for p, lines in terms:
                f = ell(f, lines[idx], p)
            idx += 1
    for _ in range(2):
        for p, lines in terms:
            f = ell(f, lines[idx], p)
        idx += 1
    return f


FINAL_EXP = (P**12 - 1) // R_ORDER * (2 * X * (6 * X * X + 3 * X + 1))


def final_exponentiation_slow(f):
    return f12_pow(f, FINAL_EXP)


def _exp_by_x(a):
    r = F12_ONE
    for bit in bin(X)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


def final_exponentiation(f):
    """Easy part then the Fuentes-Castaneda et al. hard part (the chain arkworks/halo2 use);
    equals final_exponentiation_slow (tested)."""
    f1 = f12_mul(f12_conj(f), f12_inv(f))
    r = f12_mul(f12_frob(f1, 2), f1)
    nx = lambda a: f12_conj(_exp_by_x(a))  # a^(-x), valid in the cyclotomic subgroup
    y0 = nx(r)
    y1 = f12_sqr(y0)
    y2 = f12_sqr(y1)
    y3 = f12_mul(y2, y1)
    y4 = nx(y3)
    y5 = f12_sqr(y4)
    y6 = nx(y5)
    y3 = f12_conj(y3)
    y6 = f12_conj(y6)
    y7 = f12_mul(y6, y4)
    y8 = f12_mul(y7, y3)
    y9 = f12_mul(y8, y1)
    y10 = f12_mul(y8, y4)
    y11 = f12_mul(y10, r)
    y12 = f12_frob(y9, 1)
    y13 = f12_mul(y12, y11)
    y8 = f12_frob(y8, 2)
    y14 = f12_mul(y8, y13)
    y15 = f12_frob(f12_mul(f12_conj(r), y9), 3)
    return f12_mul(y15, y14)


def pairing(p, q):  # pairings.rs:760-802
    return final_exponentiation(multi_miller_loop([(p, q)]))


# ---------------------------------------------------------------- BLS predicates
def verify_min_sig(pk_g2, msg, sig_g1, dst):
    h = hash_to_g1(msg, dst)
    f = multi_miller_loop([(sig_g1, g2_neg(G2_GEN)), (h, pk_g2)])
    return final_exponentiation(f) == F12_ONE


def verify_min_pk(pk_g1, msg, sig_g2, dst):
    h = hash_to_g2(msg, dst)
    f = multi_miller_loop([(g1_neg(G1_GEN), sig_g2), (pk_g1, h)])
    return final_exponentiation(f) == F12_ONE


def aggregate_verify_min_sig(pks_g2, msgs, aggsig_g1, dst):
    pairs = [(aggsig_g1, g2_neg(G2_GEN))] + [(hash_to_g1(m, dst), pk) for pk, m in zip(pks_g2, msgs)]
    return final_exponentiation(multi_miller_loop(pairs)) == F12_ONE


# ---------------------------------------------------------------- GLV endomorphism (randomised batch verification)
# phi(x, y) = (beta x, y) is an endomorphism of y^2 = x^3 + b over Fp and of the twist over Fp2; on the order-r
# subgroups it is multiplication by a root lambda of z^2 + z + 1 mod r.  Derived here, not copied: the cube roots
# of unity are found by exponentiation and matched against a plain scalar multiplication of the generators.
def _cube_root_of_unity(m):
    g = 2
    while pow(g, (m - 1) // 3, m) == 1:
        g += 1
    return pow(g, (m - 1) // 3, m)


def glv_constants():
    """(lambda, beta for G1, beta for G2) with (beta1 x, y) = [lambda](x, y) on G1 and (beta2 x, y) = [lambda](x, y) on G2;
    lambda is the smaller of the two roots."""
    lam = _cube_root_of_unity(R_ORDER)
    lam = min(lam, lam * lam % R_ORDER)
    b = _cube_root_of_unity(P)
    q1 = g1_mul(G1_GEN, lam)
    beta1 = next(c for c in (b, b * b % P) if (c * G1_GEN[0] % P, G1_GEN[1]) == (q1[0], q1[1]))
    q2 = g2_mul(G2_GEN, lam)
    beta2 = next(c for c in (b, b * b % P) if (tuple(c * t % P for t in G2_GEN[0]), tuple(G2_GEN[1])) == (tuple(q2[0]), tuple(q2[1])))
    return lam, beta1, beta2


def rbv_scalar(seed, index, rand_bits=64):
    """per-tuple scalar of the randomised batch verification (include/bls_bn254_b200.h): a + b lambda mod r with
    a, b the first and second rand_bits/2 bits of SHA-256(seed || LE64(index))"""
    import hashlib
    dg = hashlib.sha256(bytes(seed) + index.to_bytes(8, "little")).digest()
    h = rand_bits // 16
    a, b = int.from_bytes(dg[:h], "big"), int.from_bytes(dg[h:2 * h], "big")
    if a == 0 and b == 0:
        a = 1
    return (a + b * glv_constants()[0]) % R_ORDER
