// curve.cuh -- G1 (y^2 = x^3 + 3 over Fp) and G2 (y^2 = x^3 + 3/(9+u) over Fp2) group law,
// affine conversion, psi endomorphism, cofactor clearing and the wire codecs.
//
// Role of /root/reference/src/inner_types/g1.rs:104-118 (to affine), 283-360 (codecs), 704-841
// (add / double / multiply); g2.rs:98-112, 274-388, 685-693 (clear_cofactor), 749-886,
// 938-954 (psi).  The reference uses the complete homogeneous-projective formulas of
// Renes-Costello-Batina (eprint 2015/1060, a = 0); the same complete law is used here (one
// branch-free formula for every input, identity included), written once over a field policy.
// Results are compared after affine conversion, which is representation independent.
#pragma once
#include "tower.cuh"

namespace bn {

struct F1 {  // field policy: Fp
    typedef fp T;
    static BN_INL void add(T &r, const T &a, const T &b) { fp_add(r, a, b); }
    static BN_INL void sub(T &r, const T &a, const T &b) { fp_sub(r, a, b); }
    static BN_INL void mul(T &r, const T &a, const T &b) { fp_mulx(r, a, b); }
    static BN_INL void sqr(T &r, const T &a) { fp_mulx(r, a, a); }
    static BN_INL void neg(T &r, const T &a) { fp_neg(r, a); }
    static BN_INL void mul_by_3b(T &r, const T &a) { fp_mul9(r, a); }  // fp.rs:414-416
    static BN_INL void inv(T &r, const T &a) { fp_inv(r, a); }
    static BN_INL bool is_zero(const T &a) { return fp_is_zero(a); }
    static BN_INL void set_zero(T &r) { fp_set_zero(r); }
    static BN_INL void set_one(T &r) { fp_set_one(r); }
    static BN_INL void glv_beta(T &r, const T &a) { fp_mulx(r, a, kfp(k::GLV_BETA1)); }
    static BN_INL void select(T &r, bool c, const T &a, const T &b) { fp_select(r, c, a, b); }
};
struct F2 {  // field policy: Fp2
    typedef fp2 T;
    static BN_INL void add(T &r, const T &a, const T &b) { fp2_add(r, a, b); }
    static BN_INL void sub(T &r, const T &a, const T &b) { fp2_sub(r, a, b); }
    static BN_INL void mul(T &r, const T &a, const T &b) { fp2_mul(r, a, b); }
    static BN_INL void sqr(T &r, const T &a) { fp2_sqr(r, a); }
    static BN_INL void neg(T &r, const T &a) { fp2_neg(r, a); }
    static BN_INL void mul_by_3b(T &r, const T &a) { fp2_mul(r, a, kfp2(k::B2_3)); }  // fp2.rs:411-413
    static BN_INL void inv(T &r, const T &a) { fp2_inv(r, a); }
    static BN_INL bool is_zero(const T &a) { return fp2_is_zero(a); }
    static BN_INL void set_zero(T &r) { fp2_set_zero(r); }
    static BN_INL void set_one(T &r) { fp2_set_one(r); }
    static BN_INL void glv_beta(T &r, const T &a) { fp2_mul_fp(r, a, kfp(k::GLV_BETA2)); }
    static BN_INL void select(T &r, bool c, const T &a, const T &b) { fp2_select(r, c, a, b); }
};

template <class F>
struct point {  // homogeneous projective (X : Y : Z); identity = (0 : 1 : 0)
    typename F::T x, y, z;
};
typedef point<F1> g1p;
typedef point<F2> g2p;
struct g1a { fp x, y; bool inf; };
struct g2a { fp2 x, y; bool inf; };

template <class F>
BN_INL void pt_set_identity(point<F> &r) {
    F::set_zero(r.x);
    F::set_one(r.y);
    F::set_zero(r.z);
}
template <class F>
BN_INL void pt_from_affine(point<F> &r, const typename F::T &x, const typename F::T &y, bool inf) {
    if (inf) {
        pt_set_identity(r);
        return;
    }
    r.x = x;
    r.y = y;
    F::set_one(r.z);
}

// complete addition, RCB 2015/1060 Alg. 7 (g1.rs:744-786, g2.rs:789-831): 12 products + 2 by 3b
#ifndef BN_PT
#define BN_PT BN_NOINL
#endif
template <class F>
BN_PT void pt_add(point<F> &r, const point<F> &a, const point<F> &b) {
    typename F::T t0, t1, t2, t3, t4, x3, y3, z3;
    F::mul(t0, a.x, b.x);
    F::mul(t1, a.y, b.y);
    F::mul(t2, a.z, b.z);
    F::add(t3, a.x, a.y);
    F::add(t4, b.x, b.y);
    F::mul(t3, t3, t4);
    F::add(t4, t0, t1);
    F::sub(t3, t3, t4);
    F::add(t4, a.y, a.z);
    F::add(x3, b.y, b.z);
    F::mul(t4, t4, x3);
    F::add(x3, t1, t2);
    F::sub(t4, t4, x3);
    F::add(x3, a.x, a.z);
    F::add(y3, b.x, b.z);
    F::mul(x3, x3, y3);
    F::add(y3, t0, t2);
    F::sub(y3, x3, y3);
    F::add(x3, t0, t0);
    F::add(t0, x3, t0);
    F::mul_by_3b(t2, t2);
    F::add(z3, t1, t2);
    F::sub(t1, t1, t2);
    F::mul_by_3b(y3, y3);
    F::mul(x3, t4, y3);
    F::mul(t2, t3, t1);
    F::sub(x3, t2, x3);
    F::mul(y3, y3, t0);
    F::mul(t1, t1, z3);
    F::add(y3, t1, y3);
    F::mul(t0, t0, t3);
    F::mul(z3, z3, t4);
    F::add(z3, z3, t0);
    r.x = x3;
    r.y = y3;
    r.z = z3;
}
// doubling, RCB Alg. 9 (g1.rs:789-818, g2.rs:834-863)
template <class F>
BN_PT void pt_dbl(point<F> &r, const point<F> &a) {
    typename F::T t0, t1, t2, x3, y3, z3;
    F::sqr(t0, a.y);
    F::add(z3, t0, t0);
    F::add(z3, z3, z3);
    F::add(z3, z3, z3);
    F::mul(t1, a.y, a.z);
    F::sqr(t2, a.z);
    F::mul_by_3b(t2, t2);
    F::mul(x3, t2, z3);
    F::add(y3, t0, t2);
    F::mul(z3, t1, z3);
    F::add(t1, t2, t2);
    F::add(t2, t1, t2);
    F::sub(t0, t0, t2);
    F::mul(y3, t0, y3);
    F::add(y3, x3, y3);
    F::mul(t1, a.x, a.y);
    F::mul(x3, t0, t1);
    F::add(x3, x3, x3);
    r.x = x3;
    r.y = y3;
    r.z = z3;
}
// projective -> affine; identity -> (0, 1), inf (g1.rs:104-118, 265-271)
template <class F>
BN_INL void pt_to_affine(typename F::T &x, typename F::T &y, bool &inf, const point<F> &a) {
    typename F::T zi;
    inf = F::is_zero(a.z);
    F::inv(zi, a.z);  // 0 -> 0
    F::mul(x, a.x, zi);
    F::mul(y, a.y, zi);
    if (inf) {
        F::set_zero(x);
        F::set_one(y);
    }
}
// [k]a for a SECRET k (sign, keygen: g1.rs:821-841, g2.rs:866-886 are constant time, and so is this): every one of the
// nbits steps runs one doubling AND one complete addition, and the scalar bit only picks the result through selects, so
// the instruction stream and every memory address are independent of k (the complete RCB law has no special cases).
template <class F>
BN_INL void pt_mul_ct(point<F> &r, const point<F> &a, const uint32_t *kw, int nbits) {
    point<F> acc, t, base = a;
    pt_set_identity(acc);
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        pt_dbl(acc, acc);
        pt_add(t, acc, base);
        const bool bit = ((kw[i >> 5] >> (i & 31)) & 1) != 0;
        F::select(acc.x, bit, t.x, acc.x);
        F::select(acc.y, bit, t.y, acc.y);
        F::select(acc.z, bit, t.z, acc.z);
    }
    r = acc;
}
// [k]a for a PUBLIC k, k = nwords 32-bit words little-endian, MSB first double-and-add: branches on the bits of k
// (subgroup checks, curve constants); never called with a secret
template <class F>
BN_INL void pt_mul(point<F> &r, const point<F> &a, const uint32_t *kw, int nbits) {
    point<F> acc, base = a;
    pt_set_identity(acc);
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        pt_dbl(acc, acc);
        if ((kw[i >> 5] >> (i & 31)) & 1) pt_add(acc, acc, base);
    }
    r = acc;
}

// [a + b lambda]p for nbits-bit a, b (little-endian 32-bit words), lambda the eigenvalue of the endomorphism
// phi(x, y) = (beta x, y) on the order-r subgroup (constants derived in oracle/pymodel.py::glv_constants):
// one joint MSB-first double-and-add over the table {p, phi(p), p + phi(p)} -- nbits doublings and ~3/4 nbits
// additions for a 2*nbits-bit scalar.  Used by the randomised batch verification, whose scalars are drawn in
// this form; p must lie in the order-r subgroup for the result to equal [a + b lambda mod r]p.
template <class F>
BN_INL void pt_mul_glv(point<F> &r, const point<F> &p, const uint32_t *ka, const uint32_t *kb, int nbits) {
    point<F> tab[3], acc;
    tab[0] = p;
    tab[1] = p;
    F::glv_beta(tab[1].x, p.x);
    pt_add(tab[2], tab[0], tab[1]);
    pt_set_identity(acc);
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        pt_dbl(acc, acc);
        uint32_t sel = ((ka[i >> 5] >> (i & 31)) & 1) | (((kb[i >> 5] >> (i & 31)) & 1) << 1);
        if (sel) pt_add(acc, acc, tab[sel - 1]);
    }
    r = acc;
}

// [x]a for the curve parameter x = 0x44e992b44a6909f1 with the width-4 signed-digit string the final exponentiation uses
// (k::X_WNAF4: 14 non-zero digits): 62 doublings + 13 additions after a 4-entry table (1 doubling + 3 additions), against 27
// additions for the plain binary expansion.  Same group element as pt_mul(r, a, x, 63).
template <class F>
BN_INL void pt_mul_x(point<F> &r, const point<F> &a) {
    point<F> tab[4], d2, acc, t;
    tab[0] = a;
    pt_dbl(d2, a);
#pragma unroll 1
    for (int i = 1; i < 4; i++) pt_add(tab[i], tab[i - 1], d2);   // a, 3a, 5a, 7a
    acc = tab[(k::X_WNAF4[0] - 1) >> 1];                         // the leading digit is positive
#pragma unroll 1
    for (int i = 1; i < BN254_X_WNAF_LEN; i++) {
        pt_dbl(acc, acc);
        int d = k::X_WNAF4[i];
        if (d > 0) {
            pt_add(acc, acc, tab[(d - 1) >> 1]);
        } else if (d < 0) {
            t = tab[(-d - 1) >> 1];
            F::neg(t.y, t.y);
            pt_add(acc, acc, t);
        }
    }
    r = acc;
}

// g2.rs:938-954
BN_INL void g2_psi(g2p &r, const g2p &a) {
    fp2 t;
    fp2_conj(t, a.x);
    fp2_mul(r.x, t, kfp2(k::ENDO_U));
    fp2_conj(t, a.y);
    fp2_mul(r.y, t, kfp2(k::ENDO_V));
    fp2_conj(r.z, a.z);
}
// g2.rs:685-693: [x]P + psi(3[x]P) + psi^2([x]P) + psi^3(P)
BN_INL void g2_clear_cofactor(g2p &r, const g2p &a) {
    g2p p0, p1, p2, t;
    pt_mul_x(p0, a);
    pt_dbl(t, p0);
    pt_add(t, t, p0);
    g2_psi(p1, t);
    g2_psi(t, p0);
    g2_psi(p2, t);
    pt_add(p0, p0, p1);
    pt_add(p0, p0, p2);
    g2_psi(t, a);
    g2_psi(t, t);
    g2_psi(t, t);
    pt_add(r, p0, t);
}

// ---- wire formats -----------------------------------------------------------------------------
// G1 uncompressed x || y (g1.rs:296-302); decode g1.rs:336-358: top bit of both coordinates
// masked, x == 0 -> identity, y == 1 -> infinity flag.  Returns false for non-canonical bytes.
BN_INL void g1a_to_bytes(uint8_t *out, const g1a &a) {
    fp_to_be(out, a.x);
    fp_to_be(out + 32, a.y);
}
BN_INL bool fp_from_be_masked(fp &r, const uint8_t *b) {
    fp c;
    fp_raw_from_be(c, b);
    c.v[7] &= 0x7fffffffu;
    bool ok = fp_raw_lt_p(c);
    fp_to_mont(r, c);
    return ok;
}
BN_INL bool g1a_from_bytes(g1a &a, const uint8_t *in) {
    bool okx = fp_from_be_masked(a.x, in);
    bool oky = fp_from_be_masked(a.y, in + 32);
    if (!okx) return false;
    if (fp_is_zero(a.x)) {
        fp_set_one(a.y);
        a.inf = true;
        return true;
    }
    if (!oky) return false;
    a.inf = fp_eq(a.y, kfp(k::ONE));
    return true;
}
// G2 uncompressed x.c1 || x.c0 || y.c1 || y.c0 (g2.rs:290-300); decode g2.rs:346-388
BN_INL void g2a_to_bytes(uint8_t *out, const g2a &a) {
    fp_to_be(out, a.x.c1);
    fp_to_be(out + 32, a.x.c0);
    fp_to_be(out + 64, a.y.c1);
    fp_to_be(out + 96, a.y.c0);
}
BN_INL bool g2a_from_bytes(g2a &a, const uint8_t *in) {
    bool okx = fp_from_be(a.x.c1, in) & fp_from_be(a.x.c0, in + 32);
    bool oky = fp_from_be(a.y.c1, in + 64) & fp_from_be(a.y.c0, in + 96);
    if (!okx) return false;
    if (fp2_is_zero(a.x)) {
        fp2_set_one(a.y);
        a.inf = true;
        return true;
    }
    if (!oky) return false;
    fp2 one;
    fp2_set_one(one);
    a.inf = fp2_eq(a.y, one);
    return true;
}
// g1.rs:383-391, g2.rs:409-414
BN_INL bool g1a_on_curve(const g1a &a) {
    fp l, r;
    if (a.inf) return fp_is_zero(a.x);
    fp_mul(l, a.y, a.y);
    fp_mul(r, a.x, a.x);
    fp_mul(r, r, a.x);
    fp_add(r, r, kfp(k::THREE));
    return fp_eq(l, r);
}
BN_INL bool g2a_on_curve(const g2a &a) {
    fp2 l, r;
    if (a.inf) return true;
    fp2_sqr(l, a.y);
    fp2_sqr(r, a.x);
    fp2_mul(r, r, a.x);
    fp2_add(r, r, kfp2(k::B2));
    return fp2_eq(l, r);
}

}  // namespace bn
