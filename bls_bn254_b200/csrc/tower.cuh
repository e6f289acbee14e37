// tower.cuh -- Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v), xi = 9+u.
//
// Role of /root/reference/src/inner_types/fp2.rs:95-99,161-218,377-452, fp6.rs:123-287,
// fp12.rs:120-219 -- with the "Tier B" tower (xi = 9+u, D-twist sparse placement 0/3/4, real
// Frobenius constants) that the reference's own pairing known-answer test demands
// (pairings.rs:387-479, 970-980).  Field elements are canonical after
// every operation, so any algebraically equal formula gives bit-identical wire bytes; the
// formulas below are chosen for the GPU (Karatsuba Fp2, complex squaring, sparse line products).
#pragma once
#include "fp.cuh"

namespace bn {

struct fp2 { fp c0, c1; };
// Lane-split Fp2: two adjacent lanes of a warp hold one element, the even lane c0 and the odd lane
// c1.  Every Fp2 operation then costs each lane ONE Fp operation (a product is one fused two-term
// dot product per lane), the per-thread state of a pairing halves, and twice as many warps fit the
// same memory footprint -- see DESIGN.md "two lanes per pairing".
struct fp2l { fp v; };
// the tower above Fp2 is generic over the Fp2 representation (fp2: one thread, fp2l: lane pair)
template <class T2> struct fp6_t { T2 c0, c1, c2; };
template <class T2> struct fp12_t { fp6_t<T2> c0, c1; };
typedef fp6_t<fp2> fp6;
typedef fp12_t<fp2> fp12;
typedef fp6_t<fp2l> fp6l;
typedef fp12_t<fp2l> fp12l;

BN_INL const fp2 &kfp2(const uint32_t (&a)[2][8]) { return *reinterpret_cast<const fp2 *>(&a[0][0]); }

// ------------------------------------------------------------------------------------------ Fp2
BN_INL void fp2_set_zero(fp2 &r) { fp_set_zero(r.c0); fp_set_zero(r.c1); }
BN_INL void fp2_set_one(fp2 &r) { fp_set_one(r.c0); fp_set_zero(r.c1); }
BN_INL bool fp2_is_zero(const fp2 &a) { return fp_is_zero(a.c0) & fp_is_zero(a.c1); }
BN_INL bool fp2_eq(const fp2 &a, const fp2 &b) { return fp_eq(a.c0, b.c0) & fp_eq(a.c1, b.c1); }
// The light Fp2 operations and the Fp2 products are OUT-OF-LINE functions on pointers: the tower code above them is then a
// sequence of calls, a few KB per Fp12 routine.  Measured in round 1 (profiles/r01_notes.md): inlining them (Fp6-level
// leaves, lazy-reduction or Karatsuba Fp2 products, by-value lane primitives, shared-memory accumulators) always lost to the
// larger code -- the 32 KB instruction cache is the first-order constraint -- so those build options are gone; round 2
// re-measured two more shapes (paired products for ILP: -23 %; the slot machine of vm.cuh) with the same outcome.
#define BN_LIGHT BN_NOINL
#define BN_FP2MUL BN_NOINL
#define BN_L_ADD BN_LIGHT
#define BN_L_SUB BN_LIGHT
#define BN_L_NEG BN_LIGHT
#define BN_L_DBL BN_LIGHT
#define BN_L_CONJ BN_LIGHT
#define BN_L_XI BN_LIGHT
// Each of these first copies its operands into registers as whole (16-byte aligned) structs: all the
// 128-bit loads issue together, before the carry chains, and the result leaves as 128-bit stores.
BN_L_ADD void fp2_add(fp2 &r, const fp2 &a, const fp2 &b) {
    fp2 x = a, y = b, z;
    fp_add(z.c0, x.c0, y.c0);
    fp_add(z.c1, x.c1, y.c1);
    r = z;
}
BN_L_SUB void fp2_sub(fp2 &r, const fp2 &a, const fp2 &b) {
    fp2 x = a, y = b, z;
    fp_sub(z.c0, x.c0, y.c0);
    fp_sub(z.c1, x.c1, y.c1);
    r = z;
}
BN_L_NEG void fp2_neg(fp2 &r, const fp2 &a) {
    fp2 x = a, z;
    fp_neg(z.c0, x.c0);
    fp_neg(z.c1, x.c1);
    r = z;
}
BN_L_DBL void fp2_dbl(fp2 &r, const fp2 &a) {
    fp2 x = a, z;
    fp_dbl(z.c0, x.c0);
    fp_dbl(z.c1, x.c1);
    r = z;
}
BN_L_CONJ void fp2_conj(fp2 &r, const fp2 &a) {  // fp2.rs:428-437
    fp2 x = a, z;
    z.c0 = x.c0;
    fp_neg(z.c1, x.c1);
    r = z;
}
BN_INL void fp2_select(fp2 &r, bool c, const fp2 &a, const fp2 &b) { fp_select(r.c0, c, a.c0, b.c0); fp_select(r.c1, c, a.c1, b.c1); }
// Fused light operations (-DBN_NO_FUSED_LIGHT falls back to the two- and three-call forms): the call, the 128-bit
// operand loads and the result store are what these tiny functions cost, so a - b - c and 2(t -+ z) + t as ONE call
// each halve that overhead where the tower strings additions together (Karatsuba glue, Granger-Scott squaring).
BN_LIGHT void fp2_sub2(fp2 &r, const fp2 &a, const fp2 &b, const fp2 &c) {  // a - b - c
    fp2 x = a, y = b, w = c, z;
    fp_sub(z.c0, x.c0, y.c0); fp_sub(z.c0, z.c0, w.c0);
    fp_sub(z.c1, x.c1, y.c1); fp_sub(z.c1, z.c1, w.c1);
    r = z;
}
// one thread per element: the three-call forms stay (a fused 16-limb version made the final exponentiation 10 % slower)
BN_INL void fp2_tsub(fp2 &z, const fp2 &t) { fp2_sub(z, t, z); fp2_dbl(z, z); fp2_add(z, z, t); }  // z <- 3t - 2z
BN_INL void fp2_tadd(fp2 &z, const fp2 &t) { fp2_add(z, t, z); fp2_dbl(z, z); fp2_add(z, z, t); }  // z <- 3t + 2z

// fp2.rs:377-390 (there: 4 products).  Karatsuba with lazy reduction: three 512-bit products, the
// cross terms combined as integers, two Montgomery reductions.
//   c1 = (a0+a1)(b0+b1) - a0 b0 - a1 b1   (= a0 b1 + a1 b0 >= 0, < 2p^2 < p*2^256)
//   c0 = a0 b0 - a1 b1                     (+ p*2^256 when negative, so 0 <= . < p*2^256)
// c0 = a0 b0 + a1 (-b1), c1 = a0 b1 + a1 b0: two fused dot products (fp_dot2), no Karatsuba additions
BN_FP2MUL void fp2_mul(fp2 &r, const fp2 &a, const fp2 &b) {
    fp2 x = a, y = b, z;
    fp nb1;
    fp_neg_rs(nb1, y.c1);   // row-scanned operand of the dot product: p - b1 will do
    fp_dot2x2_inl<true>(z.c0, z.c1, x.c0, x.c1, y.c0, nb1, y.c1, y.c0);   // every factor <= p: one closing subtraction
    r = z;
}
// fp2.rs:392-402
BN_FP2MUL void fp2_sqr(fp2 &r, const fp2 &a) {
    fp2 x = a, z;
    fp s, d, m;
    fp_add_nr(s, x.c0, x.c1);   // both below 2p and used once, as the factors of the next product
    fp_sub_nr(d, x.c0, x.c1);
    fp_mul(m, x.c0, x.c1);
    fp_mul(z.c0, s, d);
    fp_dbl(z.c1, m);
    r = z;
}
BN_FP2MUL void fp2_mul_fp(fp2 &r, const fp2 &a, const fp &s) {
    fp2 x = a, z;
    fp y = s;
    fp_mul(z.c0, x.c0, y);
    fp_mul(z.c1, x.c1, y);
    r = z;
}
// (9+u)(a0 + a1 u) = (9 a0 - a1) + (a0 + 9 a1) u
BN_L_XI void fp2_mul_xi(fp2 &r, const fp2 &a) {
    fp2 x = a, z;
    fp t0, t1;
    fp_mul9(t0, x.c0);
    fp_mul9(t1, x.c1);
    fp_sub(z.c0, t0, x.c1);
    fp_add(z.c1, t1, x.c0);
    r = z;
}
BN_INL void fp2_xi_add(fp2 &r, const fp2 &a, const fp2 &b) {  // xi a + b
    fp2 t;
    fp2_mul_xi(t, a);
    fp2_add(r, t, b);
}
// fp2.rs:161-166
BN_NOINL void fp2_inv(fp2 &r, const fp2 &a) {
    fp t0, t1;
    fp_mul(t0, a.c0, a.c0);
    fp_mul(t1, a.c1, a.c1);
    fp_add(t0, t0, t1);
    fp_inv(t0, t0);
    fp_mul(r.c0, a.c0, t0);
    fp_neg(t0, t0);
    fp_mul(r.c1, a.c1, t0);
}
BN_INL uint32_t fp2_sgn0(const fp2 &a) {  // fp2.rs:95-99
    return fp_parity(a.c0) | ((uint32_t)fp_is_zero(a.c0) & fp_parity(a.c1));
}
BN_INL void fp2_norm(fp &n, const fp2 &a) {
    fp t;
    fp_mul(n, a.c0, a.c0);
    fp_mul(t, a.c1, a.c1);
    fp_add(n, n, t);
}
BN_INL bool fp2_is_square(const fp2 &a) {  // fp2.rs:441-452
    fp n;
    fp2_norm(n, a);
    return fp_is_square(n);
}
// Square root in Fp2 (role of fp2.rs:172-218).  The reference runs Alg. 9 of eprint 2012/685
// (two Fp2 exponentiations); its callers re-normalise the sign afterwards (fp2.rs:282-283,
// g2.rs:331), so either root is acceptable and the "complex" method is used here -- two Fp
// exponentiations and no inversion:
//   s = sqrt(a0^2 + a1^2);  d = (a0 + s)/2;  w = d^((p-3)/4), t = w d   (t^2 = +-d, w t = +-1)
//   d a square:      x = t + (a1/2) w u            d a non-square:  x = -(a1/2) w + t u
// Returns false when a is not a square.
BN_NOINL bool fp2_sqrt(fp2 &r, const fp2 &a) {
    fp n, s, d, w, t, h, chk;
    const fp &half = kfp(k::SVDW1_C2);  // (p-1)/2 = -1/2
    fp2_norm(n, a);
    bool ok = fp_sqrt(s, n);
    fp_add(d, a.c0, s);
    fp_mul(d, d, half);
    fp_neg(d, d);  // (a0 + s)/2
    if (fp_is_zero(d)) {  // a0 = -s: use the other root of the norm (d' = (a0 - s)/2 = a0)
        fp_sub(d, a.c0, s);
        fp_mul(d, d, half);
        fp_neg(d, d);
    }
    fp_pow_p_minus_3_div_4(w, d);
    fp_mul(t, w, d);
    fp_mul(chk, t, t);
    bool dsq = fp_eq(chk, d);
    fp_mul(h, a.c1, half);
    fp_mul(h, h, w);  // -(a1/2) w
    fp2 x, y;
    fp_neg(x.c1, h);
    x.c0 = t;
    y.c0 = h;
    y.c1 = t;
    fp2_select(r, dsq, x, y);
    if (fp_is_zero(d)) fp2_set_zero(r);  // a = 0
    return ok;
}

// constants are stored as one-thread fp2 values
BN_INL void fp2_load(fp2 &r, const fp2 &m) { r = m; }
BN_INL void fp2_store(fp2 &m, const fp2 &a) { m = a; }

// ---------------------------------------------------------------------------------- lane-split Fp2
#if BN_PTX
BN_INL uint32_t lane_par() { return threadIdx.x & 1u; }
// Shuffle mask of the lane-pair kernels.  Default: the whole warp.  Those kernels keep every warp convergent --
// whole warps exit together (BN_L2_PROLOGUE), exponents and loop schedules are the same for every item, and a pair
// that contributes nothing (identity / undecodable input) multiplies by a neutral line instead of branching around
// its Miller loop (pairing.cuh: ell) -- so a full-mask SHFL is legal and compiles to one instruction.  (Naming only the
// two lanes of a pair made ptxas wrap most shuffles in BSSY / WARPSYNC / ENDCOLLECTIVE / BSYNC: 5 instructions per
// exchanged word, 80 of the 186 instructions of the lane-pair Fp2 squaring.)
#define BN_LANES_CONVERGENT 1
BN_INL uint32_t lane_mask() { return 0xffffffffu; }
BN_INL void lane_xchg(fp &out, const fp &in) {  // out = the partner lane's `in`
    uint32_t m = lane_mask();
#pragma unroll
    for (int i = 0; i < 8; i++) out.v[i] = __shfl_xor_sync(m, in.v[i], 1);
}
BN_INL bool lane_both(bool x) {
    int o = __shfl_xor_sync(lane_mask(), (int)x, 1);
    return x & (o != 0);
}
#else
#define BN_LANES_CONVERGENT 1
// host emulation hooks, provided by tests/hostsim (two host threads per lane pair)
extern "C" uint32_t bn_host_lane_par();
extern "C" void bn_host_lane_xchg(uint32_t *out, const uint32_t *in, int nwords);
BN_INL uint32_t lane_par() { return bn_host_lane_par(); }
BN_INL void lane_xchg(fp &out, const fp &in) { bn_host_lane_xchg(out.v, in.v, 8); }
BN_INL bool lane_both(bool x) {
    uint32_t a = x, b;
    bn_host_lane_xchg(&b, &a, 1);
    return x & (b != 0);
}
#endif
// Both halves are loaded and one is selected: `lane_par() ? m.c1 : m.c0` compiles to a lane-dependent branch, and with
// that branch the fixed-line Miller kernel produced wrong results on sm_100a (nvcc 12.9) once the shuffles used the
// full mask -- also with a __syncwarp() before every exchange, so not a convergence matter (profiles/r01_notes.md).
BN_INL void fp2_load(fp2l &r, const fp2 &m) {
    fp a = m.c0, b = m.c1;
    fp_select(r.v, lane_par() != 0, b, a);
}
BN_INL void fp2_select(fp2l &r, bool c, const fp2l &a, const fp2l &b) { fp_select(r.v, c, a.v, b.v); }
BN_INL void fp2_store(fp2 &m, const fp2l &a) {
    fp *dst = lane_par() ? &m.c1 : &m.c0;   // address select, no branch
    *dst = a.v;
}
BN_INL void fp2_set_zero(fp2l &r) { fp_set_zero(r.v); }
BN_INL void fp2_set_one(fp2l &r) {
    fp z;
    fp_set_zero(z);
    fp_select(r.v, lane_par() != 0, z, kfp(k::ONE));
}
BN_INL bool fp2_is_zero(const fp2l &a) { return lane_both(fp_is_zero(a.v)); }
BN_INL bool fp2_eq(const fp2l &a, const fp2l &b) { return lane_both(fp_eq(a.v, b.v)); }
// The tower calls out-of-line functions on pointers (~8 instructions per call site, operands loaded as 128-bit vectors
// first), which keeps the lane-pair kernels as small as the one-thread ones.
BN_NOINL void fp2_add(fp2l &r, const fp2l &a, const fp2l &b) {
    fp x = a.v, y = b.v, z;
    fp_add(z, x, y);
    r.v = z;
}
BN_NOINL void fp2_sub(fp2l &r, const fp2l &a, const fp2l &b) {
    fp x = a.v, y = b.v, z;
    fp_sub(z, x, y);
    r.v = z;
}
BN_NOINL void fp2_sub2(fp2l &r, const fp2l &a, const fp2l &b, const fp2l &c) {
    fp x = a.v, y = b.v, w = c.v, z;
    fp_sub(z, x, y);
    fp_sub(z, z, w);
    r.v = z;
}
BN_NOINL void fp2_tsub(fp2l &z, const fp2l &t) {
    fp x = t.v, y = z.v, o;
    fp_sub(o, x, y);
    fp_dbl(o, o);
    fp_add(o, o, x);
    z.v = o;
}
BN_NOINL void fp2_tadd(fp2l &z, const fp2l &t) {
    fp x = t.v, y = z.v, o;
    fp_add(o, x, y);
    fp_dbl(o, o);
    fp_add(o, o, x);
    z.v = o;
}
BN_NOINL void fp2_neg(fp2l &r, const fp2l &a) {
    fp x = a.v, z;
    fp_neg(z, x);
    r.v = z;
}
BN_INL void fp2_dbl(fp2l &r, const fp2l &a) { fp2_add(r, a, a); }
BN_NOINL void fp2_conj(fp2l &r, const fp2l &a) {
    fp x = a.v, n, z;
    fp_neg(n, x);
    fp_select(z, lane_par() != 0, n, x);
    r.v = z;
}
// even lane: a0 b0 + a1 (-b1) = mine_a mine_b + theirs_a (-theirs_b)
// odd lane:  a1 b0 + a0 b1    = mine_a theirs_b + theirs_a mine_b
BN_NOINL void fp2_mul(fp2l &r, const fp2l &a, const fp2l &b) {
    fp x0 = a.v, y0 = b.v, pa, pb, npb, u, w, o;
    bool odd = lane_par() != 0;
    lane_xchg(pa, x0);
    lane_xchg(pb, y0);
    fp_neg_rs(npb, pb);     // row-scanned operand of the dot product: p - b1 will do
    fp_select(u, odd, pb, y0);
    fp_select(w, odd, y0, npb);
    fp_dot2_inl<true>(o, x0, u, pa, w);   // every factor <= p: one closing subtraction
    r.v = o;
}
BN_NOINL void fp2_sqr(fp2l &r, const fp2l &a) {
    fp x0 = a.v, pa, s, d, x, y, o;
    bool odd = lane_par() != 0;
    lane_xchg(pa, x0);
    fp_select(s, odd, pa, x0);  // a0
    fp_add_nr(x, s, pa);        // even: a0 + a1, odd: a0 + a0     (below 2p, not reduced: the factors of ONE product)
    fp_sub_nr(d, x0, pa);       // a0 - a1 + p
    fp_select(y, odd, x0, d);   // even: a0 - a1, odd: a1
    fp_mul(o, x, y);
    r.v = o;
}
BN_NOINL void fp2_mul_fp(fp2l &r, const fp2l &a, const fp &s) {
    fp x = a.v, y = s, o;
    fp_mul(o, x, y);
    r.v = o;
}
BN_NOINL void fp2_mul_xi(fp2l &r, const fp2l &a) {
    fp x0 = a.v, pa, n, t, u, o;
    lane_xchg(pa, x0);
    fp_neg_rs(n, pa);   // p - a1 in (0, p]: t + n < 2p, the one conditional subtraction of fp_add still lands in [0, p)
    fp_mul9(t, x0);
    fp_select(u, lane_par() != 0, pa, n);
    fp_add(o, t, u);
    r.v = o;
}
BN_NOINL void fp2_xi_add(fp2l &r, const fp2l &a, const fp2l &b) {  // xi a + b in one call
    fp x0 = a.v, y = b.v, pa, n, t, u, o;
    lane_xchg(pa, x0);
    fp_neg_rs(n, pa);   // as in fp2_mul_xi
    fp_mul9(t, x0);
    fp_select(u, lane_par() != 0, pa, n);
    fp_add(o, t, u);
    fp_add(o, o, y);
    r.v = o;
}
// fp2.rs:161-166 on a lane pair (both lanes run the same Fp inversion on the shared norm)
BN_INL void fp2_inv(fp2l &r, const fp2l &a) {
    fp sq, psq, n, m;
    fp_mul(sq, a.v, a.v);
    lane_xchg(psq, sq);
    fp_add(n, sq, psq);
    fp_inv(n, n);
    fp_neg(m, n);
    fp_select(n, lane_par() != 0, m, n);
    fp_mul(r.v, a.v, n);
}

// generic (unfused) forms of the fused primitives: used by the lane-split type, and by fp2 under BN_NO_FUSED
template <class T2>
BN_INL void fp2_mul_sums_g(T2 &r, const T2 &a0, const T2 &a1, const T2 &b0, const T2 &b1) {
    T2 s, u;
    fp2_add(s, a0, a1);
    fp2_add(u, b0, b1);
    fp2_mul(r, s, u);
}
template <class T2>
BN_INL void fp2_mul_sum1_g(T2 &r, const T2 &a0, const T2 &a1, const T2 &b) {
    T2 s;
    fp2_add(s, a0, a1);
    fp2_mul(r, s, b);
}
BN_INL void fp2_mul_sums(fp2l &r, const fp2l &a0, const fp2l &a1, const fp2l &b0, const fp2l &b1) { fp2_mul_sums_g(r, a0, a1, b0, b1); }
BN_INL void fp2_mul_sum1(fp2l &r, const fp2l &a0, const fp2l &a1, const fp2l &b) { fp2_mul_sum1_g(r, a0, a1, b); }
BN_INL void fp2_mul_sums(fp2 &r, const fp2 &a0, const fp2 &a1, const fp2 &b0, const fp2 &b1) { fp2_mul_sums_g(r, a0, a1, b0, b1); }
BN_INL void fp2_mul_sum1(fp2 &r, const fp2 &a0, const fp2 &a1, const fp2 &b) { fp2_mul_sum1_g(r, a0, a1, b); }

// ------------------------------------------------------------------------------------------ Fp6
template <class T2>
BN_INL void fp6_add(fp6_t<T2> &r, const fp6_t<T2> &a, const fp6_t<T2> &b) { fp2_add(r.c0, a.c0, b.c0); fp2_add(r.c1, a.c1, b.c1); fp2_add(r.c2, a.c2, b.c2); }
template <class T2>
BN_INL void fp6_sub(fp6_t<T2> &r, const fp6_t<T2> &a, const fp6_t<T2> &b) { fp2_sub(r.c0, a.c0, b.c0); fp2_sub(r.c1, a.c1, b.c1); fp2_sub(r.c2, a.c2, b.c2); }
template <class T2>
BN_INL void fp6_neg(fp6_t<T2> &r, const fp6_t<T2> &a) { fp2_neg(r.c0, a.c0); fp2_neg(r.c1, a.c1); fp2_neg(r.c2, a.c2); }
template <class T2>
BN_INL void fp6_dbl(fp6_t<T2> &r, const fp6_t<T2> &a) { fp2_dbl(r.c0, a.c0); fp2_dbl(r.c1, a.c1); fp2_dbl(r.c2, a.c2); }
// Karatsuba, 6 Fp2 products (role of fp6.rs:225-242).  r must NOT alias a or b (results are written
// as they are produced, no staging copies: code size is what bounds these kernels).
template <class T2>
BN_NOINL void fp6_mul(fp6_t<T2> &r, const fp6_t<T2> &a, const fp6_t<T2> &b) {
    T2 v0, v1, v2, t0;
    fp2_mul(v0, a.c0, b.c0);
    fp2_mul(v1, a.c1, b.c1);
    fp2_mul(v2, a.c2, b.c2);
    fp2_mul_sums(t0, a.c1, a.c2, b.c1, b.c2);
    fp2_sub2(t0, t0, v1, v2); fp2_xi_add(r.c0, t0, v0);
    fp2_mul_sums(t0, a.c0, a.c1, b.c0, b.c1);
    fp2_sub2(t0, t0, v0, v1); fp2_xi_add(r.c1, v2, t0);
    fp2_mul_sums(t0, a.c0, a.c2, b.c0, b.c2);
    fp2_sub2(t0, t0, v0, v2); fp2_add(r.c2, t0, v1);
}
// a * (b0 + b1 v): 5 Fp2 products (role of fp6.rs:123-144 mul_by_01).  r must not alias a.
template <class T2>
BN_NOINL void fp6_mul_by_01(fp6_t<T2> &r, const fp6_t<T2> &a, const T2 &b0, const T2 &b1) {
    T2 v0, v1, t0;
    fp2_mul(v0, a.c0, b0);
    fp2_mul(v1, a.c1, b1);
    fp2_mul_sum1(t0, a.c1, a.c2, b1); fp2_sub(t0, t0, v1); fp2_xi_add(r.c0, t0, v0);
    fp2_mul_sums(t0, a.c0, a.c1, b0, b1); fp2_sub2(r.c1, t0, v0, v1);
    fp2_mul_sum1(t0, a.c0, a.c2, b0); fp2_sub(t0, t0, v0); fp2_add(r.c2, t0, v1);
}
template <class T2>
BN_INL void fp6_mul_fp2(fp6_t<T2> &r, const fp6_t<T2> &a, const T2 &b) { fp2_mul(r.c0, a.c0, b); fp2_mul(r.c1, a.c1, b); fp2_mul(r.c2, a.c2, b); }
// r = a + v b,  r = a - v b  (v (b0, b1, b2) = (xi b2, b0, b1)); r must not alias b
template <class T2>
BN_INL void fp6_add_mulv(fp6_t<T2> &r, const fp6_t<T2> &a, const fp6_t<T2> &b) {
    fp2_add(r.c2, a.c2, b.c1);
    fp2_add(r.c1, a.c1, b.c0);
    fp2_xi_add(r.c0, b.c2, a.c0);
}
template <class T2>
BN_INL void fp6_sub_mulv(fp6_t<T2> &r, const fp6_t<T2> &a, const fp6_t<T2> &b) {
    T2 t;
    fp2_mul_xi(t, b.c2);
    fp2_sub(r.c2, a.c2, b.c1);
    fp2_sub(r.c1, a.c1, b.c0);
    fp2_sub(r.c0, a.c0, t);
}
// multiply by v (fp6.rs:146-152 with xi = 9+u)
template <class T2>
BN_INL void fp6_mul_v(fp6_t<T2> &r, const fp6_t<T2> &a) {
    T2 t;
    fp2_mul_xi(t, a.c2);
    r.c2 = a.c1;
    r.c1 = a.c0;
    r.c0 = t;
}
// fp6.rs:261-287
template <class T2>
BN_NOINL void fp6_inv(fp6_t<T2> &r, const fp6_t<T2> &a) {
    T2 t0, t1, t2, d, s;
    fp2_sqr(t0, a.c0); fp2_mul(s, a.c1, a.c2); fp2_mul_xi(s, s); fp2_sub(t0, t0, s);
    fp2_sqr(t1, a.c2); fp2_mul_xi(t1, t1); fp2_mul(s, a.c0, a.c1); fp2_sub(t1, t1, s);
    fp2_sqr(t2, a.c1); fp2_mul(s, a.c0, a.c2); fp2_sub(t2, t2, s);
    fp2_mul(d, a.c2, t1); fp2_mul(s, a.c1, t2); fp2_add(d, d, s); fp2_mul_xi(d, d);
    fp2_mul(s, a.c0, t0); fp2_add(d, d, s);
    fp2_inv(d, d);
    fp2_mul(r.c0, t0, d); fp2_mul(r.c1, t1, d); fp2_mul(r.c2, t2, d);
}

// ------------------------------------------------------------------------------------------ Fp12
template <class T2>
BN_INL void fp12_set_one(fp12_t<T2> &r) {
    fp2_set_one(r.c0.c0); fp2_set_zero(r.c0.c1); fp2_set_zero(r.c0.c2);
    fp2_set_zero(r.c1.c0); fp2_set_zero(r.c1.c1); fp2_set_zero(r.c1.c2);
}
template <class T2>
BN_INL bool fp12_is_one(const fp12_t<T2> &a) {
    T2 one;
    fp2_set_one(one);
    return fp2_eq(a.c0.c0, one) & fp2_is_zero(a.c0.c1) & fp2_is_zero(a.c0.c2) & fp2_is_zero(a.c1.c0) & fp2_is_zero(a.c1.c1) &
           fp2_is_zero(a.c1.c2);
}
// fp12.rs:203-210 (Karatsuba over Fp6; r may alias a or b)
template <class T2>
BN_NOINL void fp12_mul(fp12_t<T2> &r, const fp12_t<T2> &a, const fp12_t<T2> &b) {
    fp6_t<T2> s, t, u;
    fp6_add(s, a.c0, a.c1);
    fp6_add(t, b.c0, b.c1);
    fp6_mul(u, s, t);            // (a0 + a1)(b0 + b1)
    fp6_mul(s, a.c0, b.c0);      // aa
    fp6_mul(t, a.c1, b.c1);      // bb
    fp6_sub(u, u, s);
    fp6_sub(r.c1, u, t);
    fp6_add_mulv(r.c0, s, t);    // aa + v bb
}
// fp12.rs:170-180 (complex squaring, 2 Fp6 products; r may alias a)
template <class T2>
BN_NOINL void fp12_sqr(fp12_t<T2> &r, const fp12_t<T2> &a) {
    fp6_t<T2> ab, s, t;
    fp6_add(s, a.c0, a.c1);
    fp6_add_mulv(t, a.c0, a.c1);     // a0 + v a1
    fp6_mul(ab, a.c0, a.c1);
    fp6_dbl(r.c1, ab);
    fp6_mul(r.c0, s, t);             // (a0 + a1)(a0 + v a1)   [a is dead from here]
    fp6_sub(r.c0, r.c0, ab);
    fp6_sub_mulv(r.c0, r.c0, ab);    // - ab - v ab
}
template <class T2>
BN_INL void fp12_conj(fp12_t<T2> &r, const fp12_t<T2> &a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }  // fp12.rs:131-136
// fp12.rs:212-219
template <class T2>
BN_NOINL void fp12_inv(fp12_t<T2> &r, const fp12_t<T2> &a) {
    fp6_t<T2> t0, t1, t2;
    fp6_mul(t0, a.c0, a.c0);
    fp6_mul(t1, a.c1, a.c1);
    fp6_sub_mulv(t0, t0, t1);
    fp6_inv(t2, t0);
    fp6_mul(t0, a.c0, t2);
    fp6_mul(t1, a.c1, t2);
    r.c0 = t0;
    fp6_neg(r.c1, t1);
}
// f * (c0 + c3 w + c4 w^3): the D-twist sparse line product, 13 Fp2 products
// (role of fp12.rs:120-129 mul_by_014 with the placement corrected to 0/3/4)
template <class T2>
BN_NOINL void fp12_mul_by_034(fp12_t<T2> &f, const T2 &c0, const T2 &c3, const T2 &c4) {
    fp6_t<T2> t0, t1, s;
    T2 d;
    fp2_add(d, c0, c3);
    fp6_add(s, f.c0, f.c1);
    fp6_mul_by_01(t1, s, d, c4);         // (f0 + f1)(l0 + l1)
    fp6_mul_fp2(t0, f.c0, c0);           // f0 l0
    fp6_mul_by_01(s, f.c1, c3, c4);      // f1 l1
    fp6_sub(t1, t1, t0);
    fp6_sub(f.c1, t1, s);
    fp6_add_mulv(f.c0, t0, s);
}
// f <- f * l * l' for TWO sparse elements l = a + (b + c v) w, l' = a' + (b' + c' v) w (the lines of the two pairs of one Miller
// step).  The sparse-by-sparse product first,
//     l l' = (a a' + xi c c') + b b' v + (b c' + c b') v^2  +  ((a b' + a' b) + (a c' + a' c) v) w
// (six Fp2 products with Karatsuba on the three cross terms), then f times that -- one coefficient of its w part is zero:
// 6 + 5 + 6 products.  23 Fp2 products instead of the 2 x 13 of two fp12_mul_by_034.
template <class T2>
BN_NOINL void fp12_mul_by_034_034(fp12_t<T2> &f, const T2 &a, const T2 &b, const T2 &c, const T2 &a2, const T2 &b2, const T2 &c2) {
    fp6_t<T2> g0, t1, s;
    T2 g10, g11;
    {
        T2 aa, bb, cc, t;
        fp2_mul(aa, a, a2);
        fp2_mul(bb, b, b2);
        fp2_mul(cc, c, c2);
        fp2_mul_sums(t, b, c, b2, c2); fp2_sub2(g0.c2, t, bb, cc);
        fp2_mul_sums(t, a, b, a2, b2); fp2_sub2(g10, t, aa, bb);
        fp2_mul_sums(t, a, c, a2, c2); fp2_sub2(g11, t, aa, cc);
        fp2_xi_add(g0.c0, cc, aa);
        g0.c1 = bb;
    }
    // f's own halves serve as temporaries as soon as they are consumed (the stack of the Miller kernels is what they run on)
    fp6_add(s, f.c0, f.c1);                  // f0 + f1
    fp6_mul_by_01(t1, f.c1, g10, g11);       // t1 = f1 g1            (f.c1 is free now)
    fp6_mul(f.c1, f.c0, g0);                 // t0 = f0 g0 in f.c1    (f.c0 is free now)
    fp2_add(g0.c0, g0.c0, g10);              // g0 + g1
    fp2_add(g0.c1, g0.c1, g11);
    fp6_mul(f.c0, s, g0);                    // (f0 + f1)(g0 + g1) in f.c0
    fp6_sub(s, f.c0, f.c1);
    fp6_sub(s, s, t1);                       // the new f1
    fp6_add_mulv(f.c0, f.c1, t1);            // the new f0 = t0 + v t1
    f.c1 = s;
}
// coefficient of w^i: i even -> c0.c(i/2), i odd -> c1.c((i-1)/2)
template <class T2>
BN_INL T2 &fp12_coeff(fp12_t<T2> &a, int i) {
    fp6_t<T2> &h = (i & 1) ? a.c1 : a.c0;
    return (i >> 1) == 0 ? h.c0 : ((i >> 1) == 1 ? h.c1 : h.c2);
}
// x -> x^(p^kk), kk = 1, 2, 3 (role of fp12.rs:139-168 with the real constants xi^(i(p^k-1)/6))
template <class T2>
BN_NOINL void fp12_frob(fp12_t<T2> &r, const fp12_t<T2> &a, int kk) {
    fp12_t<T2> t = a;
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        T2 &c = fp12_coeff(t, i);
        if (kk & 1) fp2_conj(c, c);
        if (i) {
            const fp2 &g = *reinterpret_cast<const fp2 *>(&k::GAMMA[kk - 1][i - 1][0][0]);
            if (kk == 2) {
                fp2_mul_fp(c, c, g.c0);
            } else {
                T2 gl;
                fp2_load(gl, g);
                fp2_mul(c, c, gl);
            }
        }
    }
    r = t;
}

// pairings.rs:52-63
template <class T2>
BN_INL void fp4_square(T2 &c0, T2 &c1, const T2 &a, const T2 &b) {
    T2 t0, t1, t2;
    fp2_sqr(t0, a);
    fp2_sqr(t1, b);
    fp2_xi_add(c0, t1, t0);
    fp2_add(t2, a, b);
    fp2_sqr(t2, t2);
    fp2_sub2(c1, t2, t0, t1);
}
// Granger-Scott squaring in the cyclotomic subgroup (pairings.rs:68-115), in place: the six
// coefficients are updated where they live (four Fp2 temporaries instead of ten).
//   (z0, z1) = (c0.c0, c1.c1), (z2, z3) = (c1.c0, c0.c2), (z4, z5) = (c0.c1, c1.c2)
template <class T2>
BN_NOINL void cyclotomic_square(fp12_t<T2> &r, const fp12_t<T2> &f) {
    if (&r != &f) r = f;
    T2 &z0 = r.c0.c0, &z4 = r.c0.c1, &z3 = r.c0.c2, &z2 = r.c1.c0, &z1 = r.c1.c1, &z5 = r.c1.c2;
    T2 t0, t1, t2, t3;
    fp4_square(t0, t1, z0, z1);
    fp2_tsub(z0, t0);   // 3 t0 - 2 z0
    fp2_tadd(z1, t1);   // 3 t1 + 2 z1
    fp4_square(t0, t1, z2, z3);
    fp4_square(t2, t3, z4, z5);
    fp2_tsub(z4, t0);
    fp2_tadd(z5, t1);
    fp2_mul_xi(t0, t3);
    fp2_tadd(z2, t0);
    fp2_tsub(z3, t2);
}

// one-thread record (HBM) <-> this thread's representation
BN_INL void fp12_load(fp12 &r, const fp12 &m) { r = m; }
BN_INL void fp12_store(fp12 &m, const fp12 &a) { m = a; }
BN_INL void fp12_load(fp12l &r, const fp12 &m) {
    fp2_load(r.c0.c0, m.c0.c0); fp2_load(r.c0.c1, m.c0.c1); fp2_load(r.c0.c2, m.c0.c2);
    fp2_load(r.c1.c0, m.c1.c0); fp2_load(r.c1.c1, m.c1.c1); fp2_load(r.c1.c2, m.c1.c2);
}
BN_INL void fp12_store(fp12 &m, const fp12l &a) {
    fp2_store(m.c0.c0, a.c0.c0); fp2_store(m.c0.c1, a.c0.c1); fp2_store(m.c0.c2, a.c0.c2);
    fp2_store(m.c1.c0, a.c1.c0); fp2_store(m.c1.c1, a.c1.c1); fp2_store(m.c1.c2, a.c1.c2);
}

// wire format: 12 x 32 B big-endian in memory order c0.c0.c0, c0.c0.c1, ... (pairings.rs:499-514)
BN_INL void fp12_to_bytes(uint8_t *out, const fp12 &a) {
    const fp *x = reinterpret_cast<const fp *>(&a);
#pragma unroll 1
    for (int i = 0; i < 12; i++) fp_to_be(out + 32 * i, x[i]);
}
BN_INL bool fp12_from_bytes(fp12 &a, const uint8_t *in) {  // pairings.rs:516-579
    fp *x = reinterpret_cast<fp *>(&a);
    bool ok = true;
#pragma unroll 1
    for (int i = 0; i < 12; i++) ok &= fp_from_be(x[i], in + 32 * i);
    return ok;
}

}  // namespace bn
