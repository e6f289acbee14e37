// items.cuh -- the per-item bodies of every kernel (one message / point / pairing per thread).
// kernels.cu wraps each of these in a __global__ grid-stride launch; tests/hostsim calls the very
// same bodies in a host loop (g++ build, carry chains emulated) to unit-test the algebra on CPU.
//
// Device-internal formats (HBM, Montgomery form, array-of-structures, one 16-byte aligned
// record per item so each thread moves whole 32-byte sectors):
//   g1m  { fp x, y }        64 B      g2m  { fp2 x, y }      128 B      fp12            384 B
//   flag byte per item:  FL_INF (point at infinity)  |  FL_BAD (undecodable / failed)
#pragma once
#include "h2c.cuh"
#include "pairing.cuh"

namespace bn {

#ifndef BN_TPB_HEAVY
#define BN_TPB_HEAVY 64
#endif
#define BN_ACC(T, name) T name

enum : uint8_t { FL_INF = 1, FL_BAD = 2 };
enum : uint32_t { VERIFY_CHECK_ON_CURVE = 1u };

struct __attribute__((aligned(16))) g1m { fp x, y; };
struct __attribute__((aligned(16))) g2m { fp2 x, y; };

// ---- decode / encode (wire <-> internal) --------------------------------------------------------
// `check` is a bit set: DEC_ON_CURVE = the curve equation must hold; DEC_NO_IDENTITY = a point that decodes as the identity
// (x == 0, or y == 1: the reference's infinity convention, g1.rs:318-324 / g2.rs:329-335) is rejected like an
// undecodable one.  The verification predicates decode public keys and signatures with DEC_NO_IDENTITY (KeyValidate of
// the IETF BLS draft): an identity key would otherwise drop out of the Miller product and accept an identity signature
// on every message.
enum : int { DEC_ON_CURVE = 1, DEC_NO_IDENTITY = 4 };
BN_INL void item_g1_decode(g1m &out, uint8_t &fl, const uint8_t *in, int check) {
    g1a a;
    bool ok = g1a_from_bytes(a, in);
    if (ok && (check & DEC_ON_CURVE)) ok = g1a_on_curve(a);
    if (ok && (check & DEC_NO_IDENTITY) && a.inf) ok = false;
    out.x = a.x;
    out.y = a.y;
    fl = ok ? (a.inf ? FL_INF : 0) : FL_BAD;
}
BN_INL void item_g2_decode(g2m &out, uint8_t &fl, const uint8_t *in, int check) {
    g2a a;
    bool ok = g2a_from_bytes(a, in);
    if (ok && (check & DEC_ON_CURVE)) ok = g2a_on_curve(a);
    if (ok && (check & DEC_NO_IDENTITY) && a.inf) ok = false;
    out.x = a.x;
    out.y = a.y;
    fl = ok ? (a.inf ? FL_INF : 0) : FL_BAD;
}
BN_INL void zero_bytes(uint8_t *p, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) p[i] = 0;
}
BN_INL void item_g1_encode(uint8_t *out, const g1m &p, uint8_t fl) {
    if (fl & FL_BAD) {
        zero_bytes(out, 64);
        return;
    }
    g1a a;
    a.x = p.x;
    a.y = p.y;
    g1a_to_bytes(out, a);
}
BN_INL void item_g2_encode(uint8_t *out, const g2m &p, uint8_t fl) {
    if (fl & FL_BAD) {
        zero_bytes(out, 128);
        return;
    }
    g2a a;
    a.x = p.x;
    a.y = p.y;
    g2a_to_bytes(out, a);
}

// ---- hash to curve ---------------------------------------------------------------------------
BN_INL void item_hash_g1(g1m &out, uint8_t &fl, const uint8_t *msg, size_t len, const uint8_t *dst, uint32_t dst_len, int count) {
    g1a a;
    bool ok = hash_to_g1(a, msg, len, dst, dst_len, count);
    out.x = a.x;
    out.y = a.y;
    fl = ok ? (a.inf ? FL_INF : 0) : FL_BAD;
}
BN_INL void item_hash_g2(g2m &out, uint8_t &fl, const uint8_t *msg, size_t len, const uint8_t *dst, uint32_t dst_len, int count) {
    g2a a;
    bool ok = hash_to_g2(a, msg, len, dst, dst_len, count);
    out.x = a.x;
    out.y = a.y;
    fl = ok ? (a.inf ? FL_INF : 0) : FL_BAD;
}

// ---- Miller loops ------------------------------------------------------------------------------
// one pair (pairing(), pairings.rs:760-802 without the final exponentiation)
BN_INL void item_miller1(fp12 &f, uint8_t &fl, const g1m &p, uint8_t flp, const g2m &q, uint8_t flq) {
    const fp *px[1] = {&p.x}, *py[1] = {&p.y};
    const fp2 *qx[1] = {&q.x}, *qy[1] = {&q.y};
    bool live[1] = {((flp | flq) & (FL_INF | FL_BAD)) == 0};
    miller_loop<1, false, fp2>(f, px, py, qx, qy, live);
    fl = ((flp | flq) & FL_BAD) ? FL_BAD : 0;
}
// two pairs, shared squaring; pair 0 is (p0, -g2) through the precomputed line table when FIXED0
template <bool FIXED0>
BN_INL void item_miller2(fp12 &f, uint8_t &fl, const g1m &p0, uint8_t flp0, const g2m *q0, uint8_t flq0, const g1m &p1, uint8_t flp1,
                         const g2m &q1, uint8_t flq1) {
    const fp *px[2] = {&p0.x, &p1.x}, *py[2] = {&p0.y, &p1.y};
    const fp2 *qx[2] = {FIXED0 ? nullptr : &q0->x, &q1.x}, *qy[2] = {FIXED0 ? nullptr : &q0->y, &q1.y};
    bool live[2] = {((flp0 | flq0) & (FL_INF | FL_BAD)) == 0, ((flp1 | flq1) & (FL_INF | FL_BAD)) == 0};
    miller_loop<2, FIXED0, fp2>(f, px, py, qx, qy, live);
    fl = ((flp0 | flq0 | flp1 | flq1) & FL_BAD) ? FL_BAD : 0;
}

// two pairs, both from precomputed lines: pair 0 = (p0, -g2), pair 1 = (p1, the prepared point behind tab1 / flq1)
BN_INL void item_miller2_prep(fp12 &f, uint8_t &fl, const g1m &p0, uint8_t flp0, const g1m &p1, uint8_t flp1, const line_t<fp2> *tab1, uint8_t flq1) {
    const bool live0 = (flp0 & (FL_INF | FL_BAD)) == 0, live1 = ((flp1 | flq1) & (FL_INF | FL_BAD)) == 0;
    miller_loop_prepared<fp2>(f, p0.x, p0.y, live0, p1.x, p1.y, tab1, live1);
    fl = ((flp0 | flp1 | flq1) & FL_BAD) ? FL_BAD : 0;
}
BN_INL void item_miller2_prep_l2(fp12 &F, uint8_t &fl, const g1m &p0, uint8_t flp0, const g1m &p1, uint8_t flp1, const line_t<fp2> *tab1, uint8_t flq1) {
    const bool live0 = (flp0 & (FL_INF | FL_BAD)) == 0, live1 = ((flp1 | flq1) & (FL_INF | FL_BAD)) == 0;
    fp c0x = p0.x, c0y = p0.y, c1x = p1.x, c1y = p1.y;   // local copies: one address space per pointer parameter of the leaves
    fp12l f;
    miller_loop_prepared<fp2l>(f, c0x, c0y, live0, c1x, c1y, tab1, live1);
    fp12_store(F, f);
    fl = ((flp0 | flp1 | flq1) & FL_BAD) ? FL_BAD : 0;
}
// G2Prepared::from for one decoded point
BN_INL void item_g2_prepare(line_t<fp2> *out, const g2m &q, uint8_t fl) {
    if (fl & (FL_BAD | FL_INF)) return;      // the table of an unusable key is never read: its tuples are dead / rejected
    g2_prepare_lines(out, q.x, q.y);
}

// ---- final exponentiation / GT ---------------------------------------------------------------
BN_INL void item_final_exp(fp12 &f) { final_exponentiation(f); }

// ---- the same bodies with TWO LANES per pairing (lane-split Fp2, tower.cuh).  Both lanes of a pair
//      call these with identical arguments; `F` is the item's one-thread record in HBM and each lane
//      reads / writes its own half of it. -------------------------------------------------------------
BN_INL void item_miller1_l2(fp12 &F, uint8_t &fl, const g1m &p, uint8_t flp, const g2m &q, uint8_t flq) {
    const fp *px[1] = {&p.x}, *py[1] = {&p.y};
    fp2l lx[1], ly[1];
    fp2_load(lx[0], q.x);
    fp2_load(ly[0], q.y);
    const fp2l *qx[1] = {&lx[0]}, *qy[1] = {&ly[0]};
    bool live[1] = {((flp | flq) & (FL_INF | FL_BAD)) == 0};
    BN_ACC(fp12l, f);
    miller_loop<1, false, fp2l>(f, px, py, qx, qy, live);
    fp12_store(F, f);
    fl = ((flp | flq) & FL_BAD) ? FL_BAD : 0;
}
template <bool FIXED0>
BN_INL void item_miller2_l2(fp12 &F, uint8_t &fl, const g1m &p0, uint8_t flp0, const g2m *q0, uint8_t flq0, const g1m &p1, uint8_t flp1,
                            const g2m &q1, uint8_t flq1) {
    const fp *px[2] = {&p0.x, &p1.x}, *py[2] = {&p0.y, &p1.y};
    fp2l lx[2], ly[2];
    if (!FIXED0) {
        fp2_load(lx[0], q0->x);
        fp2_load(ly[0], q0->y);
    }
    fp2_load(lx[1], q1.x);
    fp2_load(ly[1], q1.y);
    const fp2l *qx[2] = {&lx[0], &lx[1]}, *qy[2] = {&ly[0], &ly[1]};
    bool live[2] = {((flp0 | flq0) & (FL_INF | FL_BAD)) == 0, ((flp1 | flq1) & (FL_INF | FL_BAD)) == 0};
    BN_ACC(fp12l, f);
    miller_loop<2, FIXED0, fp2l>(f, px, py, qx, qy, live);
    fp12_store(F, f);
    fl = ((flp0 | flq0 | flp1 | flq1) & FL_BAD) ? FL_BAD : 0;
}
BN_INL void item_final_exp_l2(fp12 &F) {
    BN_ACC(fp12l, f);
    fp12_load(f, F);
    final_exponentiation(f);
    fp12_store(F, f);
}

// ---- scalar multiplication (sign / keygen / test-data side; g1.rs:821-841, g2.rs:866-886): the scalar may be a
//      secret key, so this is the constant-time ladder (curve.cuh: pt_mul_ct) ------------------------------------
BN_INL void scalar_words(uint32_t kw[8], const uint8_t *be) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t *q = be + 4 * (7 - i);
        kw[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
}
BN_INL void item_g1_mul(g1m &out, uint8_t &flo, const g1m &p, uint8_t fl, const uint8_t *scalar_be) {
    uint32_t kw[8];
    scalar_words(kw, scalar_be);
    g1p a, r;
    pt_from_affine(a, p.x, p.y, (fl & FL_INF) != 0);
    pt_mul_ct(r, a, kw, 256);
    bool inf;
    pt_to_affine(out.x, out.y, inf, r);
    flo = (fl & FL_BAD) ? FL_BAD : (inf ? FL_INF : 0);
}
BN_INL void item_g2_mul(g2m &out, uint8_t &flo, const g2m &p, uint8_t fl, const uint8_t *scalar_be) {
    uint32_t kw[8];
    scalar_words(kw, scalar_be);
    g2p a, r;
    pt_from_affine(a, p.x, p.y, (fl & FL_INF) != 0);
    pt_mul_ct(r, a, kw, 256);
    bool inf;
    pt_to_affine(out.x, out.y, inf, r);
    flo = (fl & FL_BAD) ? FL_BAD : (inf ? FL_INF : 0);
}

// ---- subgroup membership (g1.rs:852-855, g2.rs:733-736: is_torsion_free = [r]P is the identity) ----
// out = 1 for a decodable point ON THE CURVE whose order divides r (the identity included).  Callers decode with the
// on-curve check, so FL_BAD already covers "not on the curve".  The same predicate as the reference's [r]P == O,
// evaluated without the 254-bit multiplication:
//   G1: #E(Fp) = r (cofactor 1): every point of the curve is in the subgroup, nothing is left to compute.
//   G2: for BN curves  Q in G2  <=>  [x+1]Q + psi([x]Q) + psi^2([x]Q) = psi^3([2x]Q)  (Dai-Lin-Zhao-Zhou, "Fast subgroup
//       membership testings on pairing-friendly curves", 2022; the test gnark-crypto uses for this curve): one 63-bit
//       multiplication and three psi instead of 254 doublings.  Checked against [r]Q on subgroup points, random points of
//       the twist, points of the cofactor subgroup and sums of both (oracle/pymodel.py, tests/test_hostsim.py, GPU suite).
BN_INL uint8_t item_g1_torsion_free(const g1m &p, uint8_t fl) {
    if (fl & FL_BAD) return 0;
    (void)p;
    return 1;
}
// Out of line on purpose: inlined into k_g2_torsion_free (behind the wire decode) the windowed [x]Q version returned 0 for
// subgroup points on sm_100a (nvcc 12.9) while the same body inside k_g2_require_subgroup and the host build were right --
// the third context-dependent miscompile of this round (profiles/r01_notes.md); one out-of-line copy serves both kernels.
BN_NOINL uint8_t item_g2_torsion_free(const g2m &p, uint8_t fl) {
    if (fl & FL_BAD) return 0;
    g2p a, r;
    pt_from_affine(a, p.x, p.y, (fl & FL_INF) != 0);
    g2p xq, b, c, t;
    pt_mul_x(xq, a);             // [x]Q
    g2_psi(b, xq);               // psi([x]Q)
    g2_psi(c, b);                // psi^2([x]Q)
    g2_psi(t, c);                // psi^3([x]Q)
    pt_dbl(t, t);                // psi^3([2x]Q)
    pt_add(r, xq, a);            // [x+1]Q
    pt_add(r, r, b);
    pt_add(r, r, c);
    // r == t as projective points: cross-multiplied coordinates (both may be the identity)
    fp2 l, m;
    fp2_mul(l, r.x, t.z); fp2_mul(m, t.x, r.z);
    bool eq = fp2_eq(l, m);
    fp2_mul(l, r.y, t.z); fp2_mul(m, t.y, r.z);
    eq &= fp2_eq(l, m);
    bool rz = fp2_is_zero(r.z), tz = fp2_is_zero(t.z);
    return (rz || tz) ? (uint8_t)(rz && tz) : (uint8_t)eq;
}

// ---- compressed codecs (g1.rs:283-288 / 311-328, g2.rs:274-283 / 309-340) ------------------------
// Decompression picks the root whose parity (G1) / sgn0 (G2) equals the flag bit, i.e. the inverse
// of to_compressed.  (The reference's G1 from_compressed selects by is_high, g1.rs:320, which does
// not invert its own to_compressed for about half of all points)
BN_INL void item_g1_compress(uint8_t *out, const g1m &p, uint8_t fl) {
    if (fl & FL_BAD) {
        zero_bytes(out, 32);
        return;
    }
    fp_to_be(out, p.x);
    out[0] |= (uint8_t)(fp_parity(p.y) << 7);
}
BN_INL void item_g1_decompress(g1m &out, uint8_t &fl, const uint8_t *in) {
    fp x, y, gx, ny;
    uint32_t flag = in[0] >> 7;
    bool ok = fp_from_be_masked(x, in);
    if (ok && fp_is_zero(x)) {
        fp_set_zero(out.x);
        fp_set_one(out.y);
        fl = FL_INF;
        return;
    }
    fp_mul(gx, x, x);
    fp_mul(gx, gx, x);
    fp_add(gx, gx, kfp(k::THREE));
    ok &= fp_sqrt(y, gx);
    fp_neg(ny, y);
    fp_select(y, (fp_parity(y) ^ flag) != 0, ny, y);
    out.x = x;
    out.y = y;
    fl = ok ? 0 : FL_BAD;
}
BN_INL void item_g2_compress(uint8_t *out, const g2m &p, uint8_t fl) {
    if (fl & FL_BAD) {
        zero_bytes(out, 64);
        return;
    }
    fp_to_be(out, p.x.c1);
    fp_to_be(out + 32, p.x.c0);
    out[0] |= (uint8_t)(fp2_sgn0(p.y) << 7);
}
BN_INL void item_g2_decompress(g2m &out, uint8_t &fl, const uint8_t *in) {
    fp2 x, y, gx, ny;
    uint32_t flag = in[0] >> 7;
    bool ok = fp_from_be_masked(x.c1, in) & fp_from_be(x.c0, in + 32);
    if (ok && fp2_is_zero(x)) {
        fp2_set_zero(out.x);
        fp2_set_one(out.y);
        fl = FL_INF;
        return;
    }
    fp2_sqr(gx, x);
    fp2_mul(gx, gx, x);
    fp2_add(gx, gx, kfp2(k::B2));
    ok &= fp2_sqrt(y, gx);
    fp2_neg(ny, y);
    fp2_select(y, (fp2_sgn0(y) ^ flag) != 0, ny, y);
    out.x = x;
    out.y = y;
    fl = ok ? 0 : FL_BAD;
}

// ---- randomised batch verification (no counterpart in the reference) -------------
// Small-exponent test: a group of tuples is accepted when
//     e(sum_i r_i sig_i, -g2) * prod_i e(r_i H(m_i), pk_i) == 1          (min-sig; min-pk symmetric)
// with per-tuple scalars r_i the caller cannot predict.  r_i = a_i + b_i lambda mod r, where a_i and b_i are
// the first and the second nbits/2 bits (big-endian) of SHA-256(seed || LE64(i)), i the tuple's index in the
// whole call, and lambda is the eigenvalue of (x, y) -> (beta x, y).  (a, b) -> a + b lambda mod r is
// injective for a, b < 2^64 (the shortest vector of the lattice {(u, v): u + v lambda = 0 mod r} has a 127-bit
// coordinate), so r_i is uniform over 2^nbits values and the test's error bound 2^-nbits holds, while [r_i]P
// costs nbits/2 doublings (curve.cuh: pt_mul_glv).  a = b = 0 is replaced by a = 1.
struct rbv_k {
    uint32_t a[2], b[2];   // little-endian words; only the low nbits/2 bits are used
};
BN_INL void rbv_scalar(rbv_k &kk, const uint8_t *seed, uint64_t idx, int nbits) {
    sha256 s;
    sha256_init(s);
    sha256_update(s, seed, 32);
#pragma unroll 1
    for (int b = 0; b < 8; b++) sha256_put(s, (uint8_t)(idx >> (8 * b)));
    uint8_t dg[32];
    sha256_final(s, dg);
    // digest words are big-endian: nbits = 64 -> a = h0, b = h1; nbits = 128 -> a = h0 h1, b = h2 h3
    auto be32 = [&](int w) { const uint8_t *q = dg + 4 * w; return ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3]; };
    if (nbits == 64) {
        kk.a[0] = be32(0); kk.a[1] = 0;
        kk.b[0] = be32(1); kk.b[1] = 0;
    } else {
        kk.a[1] = be32(0); kk.a[0] = be32(1);
        kk.b[1] = be32(2); kk.b[0] = be32(3);
    }
    if (!(kk.a[0] | kk.a[1] | kk.b[0] | kk.b[1])) kk.a[0] = 1;
}
// p <- [r]p, affine in and out (the G1 side of the per-tuple pairing); identity / undecodable points pass through
BN_INL void item_rbv_scale_g1(g1m &p, uint8_t &fl, const rbv_k &kk, int nbits) {
    if (fl & (FL_BAD | FL_INF)) return;
    g1p a, r;
    pt_from_affine(a, p.x, p.y, false);
    pt_mul_glv(r, a, kk.a, kk.b, nbits >> 1);
    bool inf;
    pt_to_affine(p.x, p.y, inf, r);
    fl = inf ? FL_INF : 0;
}
// out <- [r]p, projective (the signature side: summed over the group before it meets a pairing)
template <class F, class M>
BN_INL void item_rbv_scale_proj(point<F> &out, const M &p, uint8_t fl, const rbv_k &kk, int nbits) {
    if (fl & (FL_BAD | FL_INF)) {
        pt_set_identity(out);
        return;
    }
    point<F> a;
    pt_from_affine(a, p.x, p.y, false);
    pt_mul_glv(out, a, kk.a, kk.b, nbits >> 1);
}

}  // namespace bn
