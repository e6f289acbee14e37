"""as_written.py -- AUDIT ORACLE (test infrastructure only; nothing in the product path imports this).

A literal transcription, in Python big integers, of the pairing the reference crate's code computes AS WRITTEN:
    pairing()               /root/reference/src/inner_types/pairings.rs:760-802
    miller_loop()           pairings.rs:862-886   (iterates X >> 1, final doubling step, conjugation: the BLS12-381 loop)
    ell()                   pairings.rs:888-899   (mul_by_014: the M-twist placement)
    doubling_step/addition  pairings.rs:901-962
    final_exponentiation    pairings.rs:50-178    (the BLS12 hard part: five exponentiations by x)
    Fp2::mul_by_nonresidue  fp2.rs:415-420        (xi = 1 + u)
    Fp6                     fp6.rs:123-287        (incl. the `self.c0 - self.c2 + self.c2` typo of square(), fp6.rs:250, and
                                                   the frobenius constants of fp6.rs:162-184)
    Fp12                    fp12.rs:120-219       (mul_by_014, frobenius constant fp12.rs:151-165)

SURVEY.md section 0.4 lists why this is NOT the BN254 optimal-ate pairing (the tower is not a field for xi = 1 + u, wrong loop,
wrong sparse placement, wrong final exponent, constants that match nothing) and section 0.5 shows that the reference's own
known-answer test (pairings.rs:970-980, `pairing(g1, g2) == Gt::generator()`) pins the INTENDED function instead -- the one
oracle/bn254_oracle.c, oracle/pymodel.py and the CUDA kernels compute ("Tier B").  This module keeps the divergence checkable:
tests/test_oracle.py::test_pairing_as_written_audit asserts that the transcription reproduces the value recorded in SURVEY.md
Appendix A and that it differs from Gt::generator().  Every function cites the lines it transcribes; nothing is "fixed".
"""
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
X = 0x44e992b44a6909f1  # inner_types.rs:28


def _w(words):  # Fp::from_words: little-endian u64 limbs, canonical
    return sum(w << (64 * i) for i, w in enumerate(words)) % P


# ---- Fp2 = Fp[u]/(u^2 + 1): fp2.rs:377-437 -------------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_conj(a): return (a[0], (-a[1]) % P)          # fp2.rs:428-437: frobenius_map = conjugate
def f2_nonres(a): return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)   # fp2.rs:415-420: multiplication by 1 + u


def f2_inv(a):  # fp2.rs:161-166
    t = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * t % P, (-a[1]) * t % P)


F2_ZERO, F2_ONE = (0, 0), (1, 0)


# ---- Fp6: fp6.rs ------------------------------------------------------------------------------------------------
def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)
def f6_nonres(a): return (f2_nonres(a[2]), a[0], a[1])             # fp6.rs:146-152 mul_by_non_residue


def f6_mul(a, b):  # fp6.rs:225-242
    aa, bb, cc = f2_mul(a[0], b[0]), f2_mul(a[1], b[1]), f2_mul(a[2], b[2])
    t0 = f2_sub(f2_sub(f2_mul(f2_add(a[1], a[2]), f2_add(b[1], b[2])), bb), cc)
    t1 = f2_sub(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1])), aa), bb)
    t2 = f2_sub(f2_sub(f2_mul(f2_add(a[0], a[2]), f2_add(b[0], b[2])), aa), cc)
    return (f2_add(aa, f2_nonres(t0)), f2_add(t1, f2_nonres(cc)), f2_add(t2, bb))


def f6_sqr(a):  # fp6.rs:244-259, typo of line 250 kept: (c0 - c2 + c2) instead of (c0 - c1 + c2)
    s0 = f2_sqr(a[0])
    s1 = f2_add(f2_mul(a[0], a[1]), f2_mul(a[0], a[1]))
    s2 = f2_sqr(f2_add(f2_sub(a[0], a[2]), a[2]))
    s3 = f2_add(f2_mul(a[1], a[2]), f2_mul(a[1], a[2]))
    s4 = f2_sqr(a[2])
    return (f2_add(s0, f2_nonres(s3)), f2_add(s1, f2_nonres(s4)), f2_sub(f2_sub(f2_add(f2_add(s1, s2), s3), s0), s4))


def f6_mul_by_1(a, c1):  # fp6.rs:123-129
    return (f2_nonres(f2_mul(a[2], c1)), f2_mul(a[0], c1), f2_mul(a[1], c1))


def f6_mul_by_01(a, c0, c1):  # fp6.rs:131-144
    aa, bb = f2_mul(a[0], c0), f2_mul(a[1], c1)
    t1 = f2_add(f2_nonres(f2_mul(a[2], c1)), aa)
    t2 = f2_sub(f2_sub(f2_mul(f2_add(c0, c1), f2_add(a[0], a[1])), aa), bb)
    t3 = f2_add(f2_mul(a[2], c0), bb)
    return (t1, t2, t3)


FROB6_C1 = (_w([0x7089552b319d465, 0xc6695f92b50a8313, 0x97e83cccd117228f, 0x1229e4ffa8ddab81]), 0)      # fp6.rs:162-172
FROB6_C2 = (_w([0xe4bbdd0c2936b629, 0xbb30f162e133bacb, 0x31a9d1b6f9645366, 0x253570bea500f8dd]), 0)     # fp6.rs:174-184


def f6_frob(a):  # fp6.rs:154-187
    return (f2_conj(a[0]), f2_mul(f2_conj(a[1]), FROB6_C1), f2_mul(f2_conj(a[2]), FROB6_C2))


def f6_inv(a):  # fp6.rs:261-287
    t0, t1, t2 = f2_sqr(a[0]), f2_sqr(a[1]), f2_sqr(a[2])
    t3, t4, t5 = f2_mul(a[0], a[1]), f2_mul(a[0], a[2]), f2_mul(a[1], a[2])
    c0 = f2_sub(t0, f2_nonres(t5))
    c1 = f2_sub(f2_nonres(t2), t3)
    c2 = f2_sub(t1, t4)
    t6 = f2_inv(f2_add(f2_mul(a[0], c0), f2_nonres(f2_add(f2_mul(a[1], c1), f2_mul(a[2], c2)))))
    return (f2_mul(c0, t6), f2_mul(c1, t6), f2_mul(c2, t6))


F6_ZERO, F6_ONE = (F2_ZERO,) * 3, (F2_ONE, F2_ZERO, F2_ZERO)


# ---- Fp12: fp12.rs ----------------------------------------------------------------------------------------------
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):  # fp12.rs:203-210
    aa, bb = f6_mul(a[0], b[0]), f6_mul(a[1], b[1])
    o = f6_add(b[1], b[0])
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a[1], a[0]), o), aa), bb)
    return (f6_add(f6_nonres(bb), aa), c1)


def f12_sqr(a):  # fp12.rs:170-180
    ab = f6_mul(a[0], a[1])
    c0c1 = f6_add(a[0], a[1])
    c0 = f6_sub(f6_sub(f6_mul(f6_add(f6_nonres(a[1]), a[0]), c0c1), ab), f6_nonres(ab))
    return (c0, f6_add(ab, ab))


def f12_conj(a): return (a[0], f6_neg(a[1]))   # fp12.rs:131-136


def f12_mul_by_014(a, c0, c1, c4):  # fp12.rs:120-129
    aa = f6_mul_by_01(a[0], c0, c1)
    bb = f6_mul_by_1(a[1], c4)
    o = f2_add(c1, c4)
    r1 = f6_sub(f6_sub(f6_mul_by_01(f6_add(a[1], a[0]), c0, o), aa), bb)
    return (f6_add(f6_nonres(bb), aa), r1)


FROB12 = (_w([0x1904d3bf02bb0667, 0x790b65d7c486c4bc, 0x95a3c6b22a7fcfc, 0x4975e0cdc3905]),
          _w([0x1a1c0cb0184ddc7, 0x1f23e0957a5bcfd, 0xca2e8bae6c4f4b8, 0x7f2f50f9a63fcce]))            # fp12.rs:151-165


def f12_frob(a):  # fp12.rs:139-168
    return (f6_frob(a[0]), f6_mul(f6_frob(a[1]), (FROB12, F2_ZERO, F2_ZERO)))


def f12_inv(a):  # fp12.rs:212-219 (uses Fp6::square, typo included)
    t = f6_inv(f6_sub(f6_sqr(a[0]), f6_nonres(f6_sqr(a[1]))))
    return (f6_mul(a[0], t), f6_mul(a[1], f6_neg(t)))


# ---- Miller loop: pairings.rs:862-962 -------------------------------------------------------------------------------
def doubling_step(r):  # pairings.rs:901-930
    x, y, z = r
    tmp0, tmp1 = f2_sqr(x), f2_sqr(y)
    tmp2 = f2_sqr(tmp1)
    tmp3 = f2_sub(f2_sub(f2_sqr(f2_add(tmp1, x)), tmp0), tmp2)
    tmp3 = f2_add(tmp3, tmp3)
    tmp4 = f2_add(f2_add(tmp0, tmp0), tmp0)
    tmp6 = f2_add(x, tmp4)
    tmp5 = f2_sqr(tmp4)
    zsq = f2_sqr(z)
    nx = f2_sub(f2_sub(tmp5, tmp3), tmp3)
    nz = f2_sub(f2_sub(f2_sqr(f2_add(z, y)), tmp1), zsq)
    ny = f2_mul(f2_sub(tmp3, nx), tmp4)
    tmp2 = f2_add(tmp2, tmp2); tmp2 = f2_add(tmp2, tmp2); tmp2 = f2_add(tmp2, tmp2)
    ny = f2_sub(ny, tmp2)
    tmp3 = f2_mul(tmp4, zsq); tmp3 = f2_add(tmp3, tmp3); tmp3 = f2_neg(tmp3)
    tmp6 = f2_sub(f2_sub(f2_sqr(tmp6), tmp0), tmp5)
    tmp1 = f2_add(tmp1, tmp1); tmp1 = f2_add(tmp1, tmp1)
    tmp6 = f2_sub(tmp6, tmp1)
    tmp0 = f2_mul(nz, zsq); tmp0 = f2_add(tmp0, tmp0)
    return (nx, ny, nz), (tmp0, tmp3, tmp6)


def addition_step(r, q):  # pairings.rs:932-962
    x, y, z = r
    qx, qy = q
    zsq, ysq = f2_sqr(z), f2_sqr(qy)
    t0 = f2_mul(zsq, qx)
    t1 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(qy, z)), ysq), zsq), zsq)
    t2 = f2_sub(t0, x)
    t3 = f2_sqr(t2)
    t4 = f2_add(t3, t3); t4 = f2_add(t4, t4)
    t5 = f2_mul(t4, t2)
    t6 = f2_sub(f2_sub(t1, y), y)
    t9 = f2_mul(t6, qx)
    t7 = f2_mul(t4, x)
    nx = f2_sub(f2_sub(f2_sub(f2_sqr(t6), t5), t7), t7)
    nz = f2_sub(f2_sub(f2_sqr(f2_add(z, t2)), zsq), t3)
    t10 = f2_add(qy, nz)
    t8 = f2_mul(f2_sub(t7, nx), t6)
    t0 = f2_mul(y, t5); t0 = f2_add(t0, t0)
    ny = f2_sub(t8, t0)
    t10 = f2_sub(f2_sqr(t10), ysq)
    t10 = f2_sub(t10, f2_sqr(nz))
    t9 = f2_sub(f2_add(t9, t9), t10)
    t10 = f2_add(nz, nz)
    t6 = f2_neg(t6)
    t1 = f2_add(t6, t6)
    return (nx, ny, nz), (t10, t1, t9)


def ell(f, coeffs, p):  # pairings.rs:888-899
    c0 = (coeffs[0][0] * p[1] % P, coeffs[0][1] * p[1] % P)
    c1 = (coeffs[1][0] * p[0] % P, coeffs[1][1] * p[0] % P)
    return f12_mul_by_014(f, coeffs[2], c1, c0)


def miller_loop(p, q):  # pairings.rs:862-886 driven by the Adder of pairings.rs:761-787
    cur = (q[0], q[1], F2_ONE)
    f = F12_ONE
    found_one = False
    for b in range(63, -1, -1):
        i = (((X >> 1) >> b) & 1) == 1
        if not found_one:
            found_one = i
            continue
        cur, co = doubling_step(cur)
        f = ell(f, co, p)
        if i:
            cur, co = addition_step(cur, q)
            f = ell(f, co, p)
        f = f12_sqr(f)
    cur, co = doubling_step(cur)
    f = ell(f, co, p)
    return f12_conj(f)


# ---- final exponentiation: pairings.rs:50-178 -------------------------------------------------------------------
def fp4_square(a, b):  # pairings.rs:52-63
    t0, t1 = f2_sqr(a), f2_sqr(b)
    c0 = f2_add(f2_nonres(t1), t0)
    t2 = f2_sub(f2_sqr(f2_add(a, b)), t0)
    return c0, f2_sub(t2, t1)


def cyclotomic_square(f):  # pairings.rs:68-115
    z0, z4, z3 = f[0]
    z2, z1, z5 = f[1]
    t0, t1 = fp4_square(z0, z1)
    z0 = f2_sub(t0, z0); z0 = f2_add(f2_add(z0, z0), t0)
    z1 = f2_add(t1, z1); z1 = f2_add(f2_add(z1, z1), t1)
    t0, t1 = fp4_square(z2, z3)
    t2, t3 = fp4_square(z4, z5)
    z4 = f2_sub(t0, z4); z4 = f2_add(f2_add(z4, z4), t0)
    z5 = f2_add(t1, z5); z5 = f2_add(f2_add(z5, z5), t1)
    t0 = f2_nonres(t3)
    z2 = f2_add(t0, z2); z2 = f2_add(f2_add(z2, z2), t0)
    z3 = f2_sub(t2, z3); z3 = f2_add(f2_add(z3, z3), t2)
    return ((z0, z4, z3), (z2, z1, z5))


def cyclotomic_exp(f):  # pairings.rs:117-134 ("cycolotomic_exp")
    tmp = F12_ONE
    found_one = False
    for b in range(63, -1, -1):
        i = ((X >> b) & 1) == 1
        if found_one:
            tmp = cyclotomic_square(tmp)
        else:
            found_one = i
        if i:
            tmp = f12_mul(tmp, f)
    return f12_conj(tmp)


def final_exponentiation(f):  # pairings.rs:136-177
    t0 = f
    for _ in range(6):
        t0 = f12_frob(t0)
    t1 = f12_inv(f)
    t2 = f12_mul(t0, t1)
    t1 = t2
    t2 = f12_frob(f12_frob(t2))
    t2 = f12_mul(t2, t1)
    t1 = f12_conj(cyclotomic_square(t2))
    t3 = cyclotomic_exp(t2)
    t4 = cyclotomic_square(t3)
    t5 = f12_mul(t1, t3)
    t1 = cyclotomic_exp(t5)
    t0 = cyclotomic_exp(t1)
    t6 = cyclotomic_exp(t0)
    t6 = f12_mul(t6, t4)
    t4 = cyclotomic_exp(t6)
    t5 = f12_conj(t5)
    t4 = f12_mul(t4, f12_mul(t5, t2))
    t5 = f12_conj(t2)
    t1 = f12_mul(t1, t2)
    t1 = f12_frob(f12_frob(f12_frob(t1)))
    t6 = f12_mul(t6, t5)
    t6 = f12_frob(t6)
    t3 = f12_mul(t3, t0)
    t3 = f12_frob(f12_frob(t3))
    t3 = f12_mul(t3, t1)
    t3 = f12_mul(t3, t6)
    return f12_mul(t3, t4)


def pairing_as_written(p, q):
    """pairings.rs:760-802 on affine non-identity inputs: p = (x, y) in Fp, q = ((x0, x1), (y0, y1)) in Fp2 (canonical integers)"""
    return final_exponentiation(miller_loop(p, q))


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r
