// pairing.cuh -- optimal-ate Miller loop and final exponentiation for BN254.
//
// Role of /root/reference/src/inner_types/pairings.rs:50-178 (final_exponentiation,
// cyclotomic_square, exp by x), 609-660 / 725-757 (G2Prepared), 760-886 (pairing,
// multi_miller_loop, miller_loop), 888-962 (ell, doubling_step, addition_step).
// doubling_step / addition_step keep the reference's formulas (eprint 2010/354 Alg. 26/27) so the
// un-exponentiated MillerLoopResult is bit-identical to the oracle's; around them sits the BN254
// loop the reference's known-answer test requires: NAF(6x+2), two Frobenius
// additions, D-twist placement 0/3/4, Fuentes-Castaneda hard part scaled by 2x(6x^2+3x+1).
#pragma once
#include "curve.cuh"

namespace bn {

// Everything here is generic over the Fp2 representation T2 (fp2: one thread per pairing,
// fp2l: a lane pair per pairing), see tower.cuh.
template <class T2> struct line_t { T2 a, b, c; };  // evaluated as (a * yP) + (b * xP) w + c w^3
template <class T2> struct g2j { T2 x, y, z; };     // Jacobian accumulator of the Miller loop

// pairings.rs:901-930
template <class T2>
BN_NOINL void doubling_step(line_t<T2> &l, g2j<T2> &r) {
    T2 tmp0, tmp1, tmp2, tmp3, tmp4, tmp5, tmp6, zsq, t;
    fp2_sqr(tmp0, r.x);
    fp2_sqr(tmp1, r.y);
    fp2_sqr(tmp2, tmp1);
    fp2_add(tmp3, tmp1, r.x); fp2_sqr(tmp3, tmp3); fp2_sub2(tmp3, tmp3, tmp0, tmp2);
    fp2_dbl(tmp3, tmp3);
    fp2_dbl(tmp4, tmp0); fp2_add(tmp4, tmp4, tmp0);
    fp2_add(tmp6, r.x, tmp4);
    fp2_sqr(tmp5, tmp4);
    fp2_sqr(zsq, r.z);
    fp2_sub2(r.x, tmp5, tmp3, tmp3);
    fp2_add(t, r.z, r.y); fp2_sqr(t, t); fp2_sub2(r.z, t, tmp1, zsq);
    fp2_sub(t, tmp3, r.x); fp2_mul(r.y, t, tmp4);
    fp2_dbl(tmp2, tmp2); fp2_dbl(tmp2, tmp2); fp2_dbl(tmp2, tmp2);
    fp2_sub(r.y, r.y, tmp2);
    fp2_mul(tmp3, tmp4, zsq); fp2_dbl(tmp3, tmp3); fp2_neg(tmp3, tmp3);
    fp2_sqr(tmp6, tmp6); fp2_sub2(tmp6, tmp6, tmp0, tmp5);
    fp2_dbl(tmp1, tmp1); fp2_dbl(tmp1, tmp1);
    fp2_sub(tmp6, tmp6, tmp1);
    fp2_mul(tmp0, r.z, zsq); fp2_dbl(tmp0, tmp0);
    l.a = tmp0;
    l.b = tmp3;
    l.c = tmp6;
}
// pairings.rs:932-962
template <class T2>
BN_NOINL void addition_step(line_t<T2> &l, g2j<T2> &r, const T2 &qx, const T2 &qy) {
    T2 zsq, ysq, t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, zt;
    fp2_sqr(zsq, r.z);
    fp2_sqr(ysq, qy);
    fp2_mul(t0, zsq, qx);
    fp2_add(t1, qy, r.z); fp2_sqr(t1, t1); fp2_sub2(t1, t1, ysq, zsq); fp2_mul(t1, t1, zsq);
    fp2_sub(t2, t0, r.x);
    fp2_sqr(t3, t2);
    fp2_dbl(t4, t3); fp2_dbl(t4, t4);
    fp2_mul(t5, t4, t2);
    fp2_sub2(t6, t1, r.y, r.y);
    fp2_mul(t9, t6, qx);
    fp2_mul(t7, t4, r.x);
    fp2_sqr(r.x, t6); fp2_sub2(r.x, r.x, t5, t7); fp2_sub(r.x, r.x, t7);
    fp2_add(r.z, r.z, t2); fp2_sqr(r.z, r.z); fp2_sub2(r.z, r.z, zsq, t3);
    fp2_add(t10, qy, r.z);
    fp2_sub(t8, t7, r.x); fp2_mul(t8, t8, t6);
    fp2_mul(t0, r.y, t5); fp2_dbl(t0, t0);
    fp2_sub(r.y, t8, t0);
    fp2_sqr(t10, t10); fp2_sub(t10, t10, ysq);
    fp2_sqr(zt, r.z);
    fp2_sub(t10, t10, zt);
    fp2_dbl(t9, t9); fp2_sub(t9, t9, t10);
    fp2_dbl(t10, r.z);
    fp2_neg(t6, t6);
    fp2_dbl(t1, t6);
    l.a = t10;
    l.b = t1;
    l.c = t9;
}
// pairings.rs:888-899 with the D-twist placement
template <class T2>
BN_INL void ell(fp12_t<T2> &f, const line_t<T2> &l, const fp &px, const fp &py) {
    T2 c0, c1;
    fp2_mul_fp(c0, l.a, py);
    fp2_mul_fp(c1, l.b, px);
    fp12_mul_by_034(f, c0, c1, l.c);
}
// The same with a `live` flag: a dead pair multiplies by the neutral line 1 + 0 w + 0 w^3 (selects, no branch), which
// keeps the warps of the lane-pair kernels convergent (tower.cuh: lane_mask).
template <class T2>
BN_INL void ell_sel(fp12_t<T2> &f, const line_t<T2> &l, const fp &px, const fp &py, bool live) {
    T2 c0, c1, c2, one, zero;
    fp2_set_one(one);
    fp2_sub(zero, one, one);
    fp2_mul_fp(c0, l.a, py);
    fp2_mul_fp(c1, l.b, px);
    fp2_select(c0, live, c0, one);
    fp2_select(c1, live, c1, zero);
    fp2_select(c2, live, l.c, zero);
    fp12_mul_by_034(f, c0, c1, c2);
}
// The evaluated line alone, (a yP, b xP, c), neutralised to 1 + 0 w + 0 w^3 for a dead pair when `sel`: what the two-pair loops
// hand to fp12_mul_by_034_034.
template <class T2>
BN_INL void line_eval(T2 &c0, T2 &c3, T2 &c4, const line_t<T2> &l, const fp &px, const fp &py, bool sel, bool live) {
    fp2_mul_fp(c0, l.a, py);
    fp2_mul_fp(c3, l.b, px);
    c4 = l.c;
    if (sel) {
        T2 one, zero;
        fp2_set_one(one);
        fp2_sub(zero, one, one);
        fp2_select(c0, live, c0, one);
        fp2_select(c3, live, c3, zero);
        fp2_select(c4, live, c4, zero);
    }
}
BN_INL void line_fixed(line_t<fp2> &l, int idx) { l = *reinterpret_cast<const line_t<fp2> *>(&k::NEG_G2_LINES[idx][0][0][0]); }
BN_INL void line_fixed(line_t<fp2l> &l, int idx) {
    const line_t<fp2> &m = *reinterpret_cast<const line_t<fp2> *>(&k::NEG_G2_LINES[idx][0][0][0]);
    fp2_load(l.a, m.a);
    fp2_load(l.b, m.b);
    fp2_load(l.c, m.c);
}
// does this Fp2 representation branch around dead pairs (one thread per item) or neutralise them (lane pairs)?
template <class T2> struct skips_dead { static const bool value = true; };
#if BN_LANES_CONVERGENT
template <> struct skips_dead<fp2l> { static const bool value = false; };
#endif
// line idx of the fixed point -g2 (precomputed table in constant memory, one-thread fp2 layout).
// One thread per pairing multiplies straight out of the table; a lane pair first picks its halves.
BN_INL void ell_fixed(fp12_t<fp2> &f, int idx, const fp &px, const fp &py, bool = true) {
    ell(f, *reinterpret_cast<const line_t<fp2> *>(&k::NEG_G2_LINES[idx][0][0][0]), px, py);
}
BN_INL void ell_fixed(fp12_t<fp2l> &f, int idx, const fp &px, const fp &py, bool live = true) {
    const line_t<fp2> &m = *reinterpret_cast<const line_t<fp2> *>(&k::NEG_G2_LINES[idx][0][0][0]);
    line_t<fp2l> l;
    fp2_load(l.a, m.a);
    fp2_load(l.b, m.b);
    fp2_load(l.c, m.c);
    if (skips_dead<fp2l>::value) ell(f, l, px, py);
    else ell_sel(f, l, px, py, live);
}

#ifndef BN_PAIR_LINES
#define BN_PAIR_LINES 1
#endif
// Product of NP Miller loops with one shared squaring per step (multi_miller_loop,
// pairings.rs:808-857).  Pair i is skipped when live[i] is false (identity inputs contribute 1,
// pairings.rs:820-835).  When FIXED0, pair 0 is (P0, -g2) and its 88 lines come from the
// precomputed table (the G2Prepared role, pairings.rs:609-660) instead of being recomputed.
// Inputs are passed by pointer (no staging copies: code size bounds these kernels).
template <int NP, bool FIXED0, class T2>
BN_INL void miller_loop(fp12_t<T2> &f, const fp *const *px, const fp *const *py, const T2 *const *qx, const T2 *const *qy, const bool *live) {
    g2j<T2> r[NP];
    T2 nqy[NP];
    line_t<T2> l;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        if (FIXED0 && i == 0) continue;
        r[i].x = *qx[i];
        r[i].y = *qy[i];
        fp2_set_one(r[i].z);
        fp2_neg(nqy[i], *qy[i]);
    }
    fp12_set_one(f);
    int idx = 0;
    const bool skip = skips_dead<T2>::value;   // compile-time: branch around dead pairs, or neutralise their lines
#pragma unroll 1
    for (int s = 0; s < BN254_ATE_NAF_LEN; s++) {
        int d = k::ATE_NAF[s];
        if (s) fp12_sqr(f, f);  // f = 1 on the first step
        // Two pairs: the two lines of a step are multiplied with each other first and f once with their product (23 Fp2
        // products instead of 26).  One thread per item keeps the per-pair path when a pair is dead (it branches around it).
        const bool paired = BN_PAIR_LINES && NP == 2 && (!skip || (live[0] && live[NP - 1]));
        if (paired) {
            T2 a0, b0, c0, a1, b1, c1;
            if (FIXED0) line_fixed(l, idx);
            else doubling_step(l, r[0]);
            line_eval(a0, b0, c0, l, *px[0], *py[0], !skip, live[0]);
            doubling_step(l, r[1]);
            line_eval(a1, b1, c1, l, *px[1], *py[1], !skip, live[1]);
            fp12_mul_by_034_034(f, a0, b0, c0, a1, b1, c1);
            idx++;
            if (d) {
                if (FIXED0) line_fixed(l, idx);
                else addition_step(l, r[0], *qx[0], d > 0 ? *qy[0] : nqy[0]);
                line_eval(a0, b0, c0, l, *px[0], *py[0], !skip, live[0]);
                addition_step(l, r[1], *qx[1], d > 0 ? *qy[1] : nqy[1]);
                line_eval(a1, b1, c1, l, *px[1], *py[1], !skip, live[1]);
                fp12_mul_by_034_034(f, a0, b0, c0, a1, b1, c1);
                idx++;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NP; i++) {
                if (skip && !live[i]) continue;
                if (FIXED0 && i == 0) {
                    ell_fixed(f, idx, *px[0], *py[0], live[0]);
                } else {
                    doubling_step(l, r[i]);
                    if (skip) ell(f, l, *px[i], *py[i]);
                    else ell_sel(f, l, *px[i], *py[i], live[i]);
                }
            }
            idx++;
            if (d) {
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    if (skip && !live[i]) continue;
                    if (FIXED0 && i == 0) {
                        ell_fixed(f, idx, *px[0], *py[0], live[0]);
                    } else {
                        addition_step(l, r[i], *qx[i], d > 0 ? *qy[i] : nqy[i]);
                        if (skip) ell(f, l, *px[i], *py[i]);
                        else ell_sel(f, l, *px[i], *py[i], live[i]);
                    }
                }
                idx++;
            }
        }
    }
    // two Frobenius additions: Q1 = psi(Q), Q2 = -psi^2(Q)
#pragma unroll
    for (int i = 0; i < NP; i++) {
        if (skip && !live[i]) continue;
        if (FIXED0 && i == 0) {
            ell_fixed(f, idx, *px[0], *py[0], live[0]);
            ell_fixed(f, idx + 1, *px[0], *py[0], live[0]);
        } else {
            T2 q1x, q1y, q2x, q2y, t, eu, ev;
            fp2_load(eu, kfp2(k::ENDO_U));
            fp2_load(ev, kfp2(k::ENDO_V));
            fp2_conj(t, *qx[i]); fp2_mul(q1x, t, eu);
            fp2_conj(t, *qy[i]); fp2_mul(q1y, t, ev);
            fp2_conj(t, q1x); fp2_mul(q2x, t, eu);
            fp2_conj(t, q1y); fp2_mul(q2y, t, ev);
            fp2_neg(q2y, q2y);
            addition_step(l, r[i], q1x, q1y);
            if (skip) ell(f, l, *px[i], *py[i]);
            else ell_sel(f, l, *px[i], *py[i], live[i]);
            addition_step(l, r[i], q2x, q2y);
            if (skip) ell(f, l, *px[i], *py[i]);
            else ell_sel(f, l, *px[i], *py[i], live[i]);
        }
    }
}

// ---- G2Prepared (pairings.rs:609-660, 725-757): the 88 line triples of a G2 point, computed once per public key ----------
// out[0..87] in consumption order: per NAF digit a doubling line [and an addition line], then the two Frobenius additions --
// the schedule miller_loop runs, with the same doubling_step / addition_step, so a loop fed from the table is bit-identical.
BN_INL void g2_prepare_lines(line_t<fp2> *out, const fp2 &qx, const fp2 &qy) {
    g2j<fp2> r;
    fp2 nqy;
    line_t<fp2> l;
    r.x = qx;
    r.y = qy;
    fp2_set_one(r.z);
    fp2_neg(nqy, qy);
    int idx = 0;
#pragma unroll 1
    for (int s = 0; s < BN254_ATE_NAF_LEN; s++) {
        int d = k::ATE_NAF[s];
        doubling_step(l, r);
        out[idx++] = l;
        if (d) {
            addition_step(l, r, qx, d > 0 ? qy : nqy);
            out[idx++] = l;
        }
    }
    fp2 q1x, q1y, q2x, q2y, t, eu, ev;
    fp2_load(eu, kfp2(k::ENDO_U));
    fp2_load(ev, kfp2(k::ENDO_V));
    fp2_conj(t, qx); fp2_mul(q1x, t, eu);
    fp2_conj(t, qy); fp2_mul(q1y, t, ev);
    fp2_conj(t, q1x); fp2_mul(q2x, t, eu);
    fp2_conj(t, q1y); fp2_mul(q2y, t, ev);
    fp2_neg(q2y, q2y);
    addition_step(l, r, q1x, q1y);
    out[idx++] = l;
    addition_step(l, r, q2x, q2y);
    out[idx++] = l;
}
// line `tab[idx]` of a prepared point (HBM, one-thread fp2 layout) -> this thread's / lane's representation
BN_INL void line_load(line_t<fp2> &l, const line_t<fp2> &m) { l = m; }
BN_INL void line_load(line_t<fp2l> &l, const line_t<fp2> &m) {
    const bool odd = lane_par() != 0;
    const fp *pa = odd ? &m.a.c1 : &m.a.c0, *pb = odd ? &m.b.c1 : &m.b.c0, *pc = odd ? &m.c.c1 : &m.c.c0;   // address selects, one load path
    l.a.v = *pa;
    l.b.v = *pb;
    l.c.v = *pc;
}
// Two-pair Miller loop with BOTH line sequences precomputed: pair 0 = (P0, -g2) from the constant table, pair 1 = (P1, Q1)
// from Q1's prepared table in HBM.  No twist-point arithmetic at all: per step one Fp12 squaring and two sparse products
// (multi_miller_loop over G2Prepared terms, pairings.rs:808-857).
template <class T2>
BN_INL void miller_loop_prepared(fp12_t<T2> &f, const fp &p0x, const fp &p0y, bool live0, const fp &p1x, const fp &p1y, const line_t<fp2> *tab1, bool live1) {
    const bool skip = skips_dead<T2>::value;
    line_t<T2> l;
    fp12_set_one(f);
    int idx = 0;
#pragma unroll 1
    for (int s = 0; s < BN254_ATE_NAF_LEN + 1; s++) {
        // s < NAF_LEN: the doubling line and, for a non-zero digit, the addition line; s == NAF_LEN: the two Frobenius lines
        const int nlines = s < BN254_ATE_NAF_LEN ? (k::ATE_NAF[s] ? 2 : 1) : 2;
        if (s && s < BN254_ATE_NAF_LEN) fp12_sqr(f, f);
#pragma unroll 1
        for (int j = 0; j < nlines; j++) {
#if BN_PAIR_LINES
            if (!skip || (live0 && live1)) {   // both lines of the step first, then f once (fp12_mul_by_034_034)
                T2 a0, b0, c0, a1, b1, c1;
                line_fixed(l, idx);
                line_eval(a0, b0, c0, l, p0x, p0y, !skip, live0);
                line_load(l, tab1[idx]);
                line_eval(a1, b1, c1, l, p1x, p1y, !skip, live1);
                fp12_mul_by_034_034(f, a0, b0, c0, a1, b1, c1);
                idx++;
                continue;
            }
#endif
            if (!skip || live0) ell_fixed(f, idx, p0x, p0y, live0);
            if (!skip || live1) {
                line_load(l, tab1[idx]);
                if (skip) ell(f, l, p1x, p1y);
                else ell_sel(f, l, p1x, p1y, live1);
            }
            idx++;
        }
    }
}

// f^x (x = 0x44e992b44a6909f1, 63 bits) in the cyclotomic subgroup, then conjugate
// (pairings.rs:117-134 role).  Width-4 signed window over f, f^3, f^5, f^7 (a negative digit uses the
// conjugate, which is the inverse here): 63 cyclotomic squarings + 16 products instead of 62 + 27.
#ifndef BN_XWIN
#define BN_XWIN 4
#endif
template <class T2>
BN_NOINL void cyclotomic_exp_neg_x(fp12_t<T2> &r, const fp12_t<T2> &f) {
#if BN_XWIN == 4
    const int NPW = 4, LEN = BN254_X_WNAF_LEN;
    const int8_t *dig = k::X_WNAF4;
#else
    const int NPW = 2, LEN = BN254_X_WNAF3_LEN;
    const int8_t *dig = k::X_WNAF3;
#endif
    fp12_t<T2> pw[NPW], t;
    pw[0] = f;
    cyclotomic_square(t, f);
#pragma unroll 1
    for (int i = 1; i < NPW; i++) fp12_mul(pw[i], pw[i - 1], t);
    t = pw[(dig[0] - 1) >> 1];
#pragma unroll 1
    for (int i = 1; i < LEN; i++) {
        cyclotomic_square(t, t);
        int d = dig[i];
        if (d > 0) {
            fp12_mul(t, t, pw[(d - 1) >> 1]);
        } else if (d < 0) {  // t * conj(x) = conj(conj(t) * x): no temporary
            fp12_conj(t, t);
            fp12_mul(t, t, pw[(-d - 1) >> 1]);
            fp12_conj(t, t);
        }
    }
    fp12_conj(r, t);
}
// pairings.rs:50-178 role: easy part f^((p^6-1)(p^2+1)), then the Fuentes-Castaneda hard part.
// `f` is both input and output; five Fp12 temporaries.
template <class T2>
BN_INL void final_exponentiation(fp12_t<T2> &f) {
    fp12_t<T2> r, t, a, b, c;
    fp12_conj(r, f);
    fp12_inv(t, f);
    fp12_mul(t, r, t);          // f^(p^6-1)
    fp12_frob(r, t, 2);
    fp12_mul(r, r, t);          // r = f^((p^6-1)(p^2+1))
    cyclotomic_exp_neg_x(a, r);     // y0
    cyclotomic_square(a, a);        // y1
    cyclotomic_square(t, a);        // y2
    fp12_mul(b, t, a);              // y3
    cyclotomic_exp_neg_x(c, b);     // y4
    cyclotomic_square(t, c);        // y5
    cyclotomic_exp_neg_x(f, t);     // y6
    fp12_conj(b, b);
    fp12_conj(f, f);
    fp12_mul(t, f, c);          // y7
    fp12_mul(b, t, b);          // y8
    fp12_mul(f, b, a);          // y9  (a = y1 free afterwards)
    fp12_mul(t, b, c);          // y10
    fp12_mul(t, t, r);          // y11
    fp12_frob(a, f, 1);         // y12
    fp12_mul(t, a, t);          // y13
    fp12_frob(b, b, 2);
    fp12_mul(t, b, t);          // y14
    fp12_conj(a, r);
    fp12_mul(a, a, f);          // y15
    fp12_frob(a, a, 3);
    fp12_mul(f, a, t);
}

}  // namespace bn
