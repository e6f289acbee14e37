// fp.cuh -- BN254 base field Fp as 8 x 32-bit Montgomery limbs (R = 2^256) for sm_100a.
//
// Replaces (role of) /root/reference/src/inner_types/fp.rs:388-412 (add/sub/neg/multiply/square),
// 207-243 (invert, sqrt), 428-431 (is_square), 115-122 (from_okm), 246-260 (to/from_repr).
// The reference stores canonical residues and reduces with crypto-bigint's const_rem_wide; here
// every value lives in Montgomery form on the device and is canonical (fully reduced, < p)
// after every operation, so conversion back at the wire gives the same bytes.
//
// Multiplication is an operand-scanning Montgomery product on two 256-bit accumulators that
// are offset by one 32-bit word ("even" lanes hold a0,a2,a4,a6 x b_i, "odd" lanes hold
// a1,a3,a5,a7 x b_i).  Each 32x32->64 product therefore lands on a 64-bit aligned lane of its
// accumulator and a whole row is ONE carry chain of mad.lo.cc / madc.hi.cc pairs, which ptxas
// turns into IMAD.WIDE.U32(.X) -- 16 wide IMADs + 1 IMAD per row, 264 IMAD-ops per product.
//
// Under nvcc every function is __device__ only.  tests/hostsim compiles the very same headers
// with g++ (carry chains emulated in C) so the algebra is unit-tested without a GPU.  That
// build is test infrastructure; libbls_bn254_b200.so contains device code only.
#pragma once
#include <stdint.h>

#include "bn254_constants.cuh"

#if defined(__CUDACC__)
#define BN_HD __device__
#define BN_INL __device__ __forceinline__
#define BN_NOINL __device__ __noinline__
#else
#define BN_HD
#define BN_INL inline __attribute__((always_inline))
#define BN_NOINL __attribute__((noinline))
#endif

#if defined(__CUDA_ARCH__)
#define BN_PTX 1
#else
#define BN_PTX 0
#endif

// Executed-work counter of the host build (tests/hostsim): every multiplier routine reports the INT32 multiply-add
// operations it issues (one IMAD.WIDE = 2), so that tools/count_ops.py can freeze the EXECUTED work per item next to the
// canonical counts of SURVEY.md 8(d) (profiles/executed_ops.json).  Compiles to nothing under nvcc.
#if !BN_PTX
extern "C" void bn_host_count_ops(unsigned ops);
#define BN_COUNT_OPS(n) bn_host_count_ops(n)
#else
#define BN_COUNT_OPS(n) ((void)0)
#endif

namespace bn {

#ifndef BN_FP_ALIGN
#define BN_FP_ALIGN 16   // 128-bit vector loads / stores of field elements (local, global and constant memory)
#endif
struct alignas(BN_FP_ALIGN) fp {
    uint32_t v[8];
};

// modulus limbs as literals (immediates in the IMAD stream)
#define BN_P0 0xd87cfd47u
#define BN_P1 0x3c208c16u
#define BN_P2 0x6871ca8du
#define BN_P3 0x97816a91u
#define BN_P4 0x8181585du
#define BN_P5 0xb85045b6u
#define BN_P6 0xe131a029u
#define BN_P7 0x30644e72u

BN_INL uint32_t p_limb(int i) {
    switch (i) {
        case 0: return BN_P0;
        case 1: return BN_P1;
        case 2: return BN_P2;
        case 3: return BN_P3;
        case 4: return BN_P4;
        case 5: return BN_P5;
        case 6: return BN_P6;
        default: return BN_P7;
    }
}

// ---------------------------------------------------------------------------------------------
// carry-chain primitives.  Device: one asm statement per chain (the CC flag never crosses a
// statement boundary).  Host: the same chains in 64-bit C.
// ---------------------------------------------------------------------------------------------

// r = a + b (256 bit), returns carry out
BN_INL uint32_t add_n(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t c;
#if BN_PTX
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}

// r = a - b (256 bit), returns 0 or 0xffffffff (all ones when a < b)
BN_INL uint32_t sub_n(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t m;
#if BN_PTX
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(m)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    m = br ? 0xffffffffu : 0u;
#endif
    return m;
}

// r = a - p, returns borrow mask (all ones when a < p)
BN_INL uint32_t sub_p(uint32_t r[8], const uint32_t a[8]) {
    uint32_t m;
#if BN_PTX
    asm("sub.cc.u32 %0, %9, 0xd87cfd47;\n\t"
        "subc.cc.u32 %1, %10, 0x3c208c16;\n\t"
        "subc.cc.u32 %2, %11, 0x6871ca8d;\n\t"
        "subc.cc.u32 %3, %12, 0x97816a91;\n\t"
        "subc.cc.u32 %4, %13, 0x8181585d;\n\t"
        "subc.cc.u32 %5, %14, 0xb85045b6;\n\t"
        "subc.cc.u32 %6, %15, 0xe131a029;\n\t"
        "subc.cc.u32 %7, %16, 0x30644e72;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(m)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - p_limb(i) - br;
        r[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    m = br ? 0xffffffffu : 0u;
#endif
    return m;
}

// r += p & mask
BN_INL void add_p_masked(uint32_t r[8], uint32_t mask) {
#if BN_PTX
    uint32_t q0 = BN_P0 & mask, q1 = BN_P1 & mask, q2 = BN_P2 & mask, q3 = BN_P3 & mask, q4 = BN_P4 & mask, q5 = BN_P5 & mask,
             q6 = BN_P6 & mask, q7 = BN_P7 & mask;
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
        : "r"(q0), "r"(q1), "r"(q2), "r"(q3), "r"(q4), "r"(q5), "r"(q6), "r"(q7));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)r[i] + (p_limb(i) & mask);
        r[i] = (uint32_t)t;
        t >>= 32;
    }
#endif
}

// acc(256 bit) = {x0,x1,x2,x3} (four 32-bit values on 64-bit lanes) * y      (no carries possible)
BN_INL void mul_lanes(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if BN_PTX
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]), "=&r"(acc[6]), "=&r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    uint64_t t;
    t = (uint64_t)x0 * y; acc[0] = (uint32_t)t; acc[1] = (uint32_t)(t >> 32);
    t = (uint64_t)x1 * y; acc[2] = (uint32_t)t; acc[3] = (uint32_t)(t >> 32);
    t = (uint64_t)x2 * y; acc[4] = (uint32_t)t; acc[5] = (uint32_t)(t >> 32);
    t = (uint64_t)x3 * y; acc[6] = (uint32_t)t; acc[7] = (uint32_t)(t >> 32);
#endif
}

#if !BN_PTX
// host model of one lane chain: acc += lanes * y + cin (cin enters at word 0); returns carry out
BN_INL uint32_t host_mad_lanes(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y, uint32_t cin) {
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t c = cin;
    for (int k = 0; k < 4; k++) {
        uint64_t prod = (uint64_t)x[k] * y;
        uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)prod + c;
        acc[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        acc[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
    return (uint32_t)c;
}
#endif

// acc += {x0..x3} * y; the carry out (weight 2^256) is added to `top`
BN_INL void mad_lanes(uint32_t acc[8], uint32_t &top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if BN_PTX
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    top += host_mad_lanes(acc, x0, x1, x2, x3, y, 0);
#endif
}

// acc += {x0..x3} * y where the caller knows the sum cannot overflow 256 bits (no carry out)
BN_INL void mad_lanes_nc(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if BN_PTX
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    (void)host_mad_lanes(acc, x0, x1, x2, x3, y, 0);
#endif
}

// e0 += x (carry c, weight of acc word 0); acc += {x0..x3} * y + c.  The chain cannot overflow
// by construction (see fp_mul) so no carry out is produced.
BN_INL void mad_lanes_cin(uint32_t acc[8], uint32_t &e0, uint32_t x, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if BN_PTX
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.u32 %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(e0)
        : "r"(x), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    uint64_t s = (uint64_t)e0 + x;
    e0 = (uint32_t)s;
    (void)host_mad_lanes(acc, x0, x1, x2, x3, y, (uint32_t)(s >> 32));
#endif
}

// acc += {p0,p2,p4,p6} * m, carry out added to top
BN_INL void mad_p_even(uint32_t acc[8], uint32_t &top, uint32_t m) {
#if BN_PTX
    asm("mad.lo.cc.u32 %0, %9, 0xd87cfd47, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, 0xd87cfd47, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, 0x6871ca8d, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, 0x6871ca8d, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, 0x8181585d, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, 0x8181585d, %5;\n\t"
        "madc.lo.cc.u32 %6, %9, 0xe131a029, %6;\n\t"
        "madc.hi.cc.u32 %7, %9, 0xe131a029, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(m));
#else
    top += host_mad_lanes(acc, BN_P0, BN_P2, BN_P4, BN_P6, m, 0);
#endif
}

// acc += {p1,p3,p5,p7} * m (cannot overflow, see fp_mul)
BN_INL void mad_p_odd(uint32_t acc[8], uint32_t m) {
#if BN_PTX
    asm("mad.lo.cc.u32 %0, %8, 0x3c208c16, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, 0x3c208c16, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, 0x97816a91, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, 0x97816a91, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, 0xb85045b6, %4;\n\t"
        "madc.hi.cc.u32 %5, %8, 0xb85045b6, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, 0x30644e72, %6;\n\t"
        "madc.hi.u32 %7, %8, 0x30644e72, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(m));
#else
    (void)host_mad_lanes(acc, BN_P1, BN_P3, BN_P5, BN_P7, m, 0);
#endif
}

// ---------------------------------------------------------------------------------------------
// field operations (all results canonical, < p)
// ---------------------------------------------------------------------------------------------
BN_INL void fp_reduce_once(fp &r) {  // r in [0, 2p) -> [0, p)
    uint32_t t[8];
    uint32_t lt = sub_p(t, r.v);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = lt ? r.v[i] : t[i];
}

BN_INL void fp_add(fp &r, const fp &a, const fp &b) {  // fp.rs:388-390
    fp s;
    (void)add_n(s.v, a.v, b.v);  // < 2p < 2^255: no carry
    fp_reduce_once(s);
    r = s;
}
BN_INL void fp_sub(fp &r, const fp &a, const fp &b) {  // fp.rs:396-398
    fp s;
    uint32_t m = sub_n(s.v, a.v, b.v);
    add_p_masked(s.v, m);
    r = s;
}
BN_INL void fp_dbl(fp &r, const fp &a) { fp_add(r, a, a); }
// Unreduced sum / difference for values that go straight into ONE product as its two factors: a + b in [0, 2p) and
// a - b + p in (0, 2p).  fp_mul of two values below 2p stays inside its frame (rows < 2p 2^32, as in fp_dot2) and ends below
// (4p^2 + R p) / R < 1.76 p, so its single conditional subtraction still returns the canonical residue.
BN_INL void fp_add_nr(fp &r, const fp &a, const fp &b) {
    fp s;
    (void)add_n(s.v, a.v, b.v);
    r = s;
}
BN_INL void fp_sub_nr(fp &r, const fp &a, const fp &b) {
    fp s;
    (void)sub_n(s.v, a.v, b.v);
    add_p_masked(s.v, 0xffffffffu);
    r = s;
}
BN_INL bool fp_is_zero(const fp &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0; }
BN_INL bool fp_eq(const fp &a, const fp &b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
BN_INL void fp_set_zero(fp &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
BN_INL void fp_neg(fp &r, const fp &a) {  // fp.rs:400-402
    fp z;
    fp_set_zero(z);
    fp_sub(r, z, a);
}
// r = p - a, in (0, p]: the negation WITHOUT the a == 0 fix-up (p instead of 0), one borrow chain against immediates instead of
// fp_neg's three.  Only for a value that goes straight into a product as the ROW-SCANNED operand (b of fp_mul, b / d of
// fp_dot2), which may be any 256-bit value congruent to what is meant: p there multiplies to 0 mod p like 0 does.
BN_INL void fp_neg_rs(fp &r, const fp &a) {
    const uint32_t pl[8] = {p_limb(0), p_limb(1), p_limb(2), p_limb(3), p_limb(4), p_limb(5), p_limb(6), p_limb(7)};
    fp s;
    (void)sub_n(s.v, pl, a.v);
    r = s;
}
BN_INL void fp_select(fp &r, bool c, const fp &a, const fp &b) {  // r = c ? a : b
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
}

// Montgomery product a*b/R mod p; inputs < p (any input < 2^256 with a*b < p*R gives a value
// congruent to a*b/R, fully reduced as long as the pre-reduction value is < 2p).
//
// Running value T = E + O * 2^32 (E: frame words 0..7, O: frame words 1..8).  Per row:
//   T += a*b_i   (even limbs of a into E, odd limbs into O)
//   m  = E[0] * (-p^-1)           (only E reaches word 0)
//   T += m*p     -> E[0] == 0
//   T >>= 32     -> E' = O + E[1],  O' = E >> 64   (pure renaming; the one-word add of E[1] is the
//                   add.cc that opens the next odd chain, which starts at exactly that weight)
// T < 2^32*2p + 2p < 2^288 throughout, so O (< T / 2^32) never overflows; E may carry out of
// word 7 (weight = O[7]) and that carry is folded into O[7] by the closing addc of the chain.
BN_INL void fp_mul_inl(fp &r, const fp &a, const fp &b) {  // fp.rs:404-407
    BN_COUNT_OPS(264);   // 8 rows x (16 + 16 wide-IMAD halves... = 128 + 128 ops) + 8 quotient digits
    uint32_t E[8], O[8], m;
    mul_lanes(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    mul_lanes(O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
    m = E[0] * BN254_INV32;
    mad_p_odd(O, m);
    mad_p_even(E, O[7], m);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        // shift by one word: E' = O (+ E[1] at word 0), O' = (E[2..7], 0, 0)
        uint32_t x = E[1];
        uint32_t N[8] = {E[2], E[3], E[4], E[5], E[6], E[7], 0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; k++) E[k] = O[k];
        mad_lanes_cin(N, E[0], x, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
        mad_lanes(E, N[7], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
        m = E[0] * BN254_INV32;
        mad_p_odd(N, m);
        mad_p_even(E, N[7], m);
#pragma unroll
        for (int k = 0; k < 8; k++) O[k] = N[k];
    }
    // final shift and merge: r = O + E[1] + (E >> 64) * 2^32      (< 2p)
    fp s;
    uint32_t hi[8] = {E[1], E[2], E[3], E[4], E[5], E[6], E[7], 0u};
    (void)add_n(s.v, O, hi);
    fp_reduce_once(s);
    r = s;
}
// Fused two-term Montgomery dot product (a*b + c*d)/R mod p, canonical result.  Same frame as
// fp_mul_inl with two product rows per step: T stays < 3p + 1 (T' = (T + a b_i + c d_i + m p)/2^32),
// T + 3p*2^32 < 2^288 so the 9-word frame holds, and two conditional subtractions finish.
// This is how Fp2 products are formed: c0 = a0 b0 + a1 (-b1), c1 = a0 b1 + a1 b0 -- 2 x 200 wide
// IMADs and one negation instead of Karatsuba's 3 x 136 and five additions.
// FACTORS_LE_P = true: the caller guarantees b, d <= p as well (a, c < p always).  The closing value is exactly
// (a b + c d + M p) / R with M < R, i.e. below p (1 + 2p/R) < 1.38 p then, and ONE conditional subtraction finishes; with
// arbitrary 256-bit row-scanned operands it is below 3p and takes two.
template <bool FACTORS_LE_P = false>
BN_INL void fp_dot2_inl(fp &r, const fp &a, const fp &b, const fp &c, const fp &d) {
    BN_COUNT_OPS(392);   // two product rows (2 x 128) + one reduction (128) + 8 quotient digits
    uint32_t E[8], O[8], m;
    mul_lanes(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    mul_lanes(O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
    mad_lanes_nc(O, c.v[1], c.v[3], c.v[5], c.v[7], d.v[0]);
    mad_lanes(E, O[7], c.v[0], c.v[2], c.v[4], c.v[6], d.v[0]);
    m = E[0] * BN254_INV32;
    mad_p_odd(O, m);
    mad_p_even(E, O[7], m);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        uint32_t x = E[1];
        uint32_t N[8] = {E[2], E[3], E[4], E[5], E[6], E[7], 0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; k++) E[k] = O[k];
        mad_lanes_cin(N, E[0], x, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
        mad_lanes_nc(N, c.v[1], c.v[3], c.v[5], c.v[7], d.v[i]);
        mad_lanes(E, N[7], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
        mad_lanes(E, N[7], c.v[0], c.v[2], c.v[4], c.v[6], d.v[i]);
        m = E[0] * BN254_INV32;
        mad_p_odd(N, m);
        mad_p_even(E, N[7], m);
#pragma unroll
        for (int k = 0; k < 8; k++) O[k] = N[k];
    }
    fp s;
    uint32_t hi[8] = {E[1], E[2], E[3], E[4], E[5], E[6], E[7], 0u};
    (void)add_n(s.v, O, hi);  // < 3p + 1 < 2^256
    fp_reduce_once(s);
    if (!FACTORS_LE_P) fp_reduce_once(s);
    r = s;
}

// Two independent fused dot products advanced row by row in lock step:
//   r0 = (a b0 + c d0)/R,  r1 = (a b1 + c d1)/R      (same multiplicands a, c: an Fp2 product)
// One warp alone reaches only ~64 % of the IMAD pipe on a single dot product (dependent carry chains,
// measured with k_fp_peak); two interleaved instruction streams give the scheduler what a second warp
// would.  Same frame and bounds as fp_dot2_inl.
template <bool FACTORS_LE_P = false>
BN_INL void fp_dot2x2_inl(fp &r0, fp &r1, const fp &a, const fp &c, const fp &b0, const fp &d0, const fp &b1, const fp &d1) {
    BN_COUNT_OPS(784);
    uint32_t E0[8], O0[8], E1[8], O1[8], m0, m1;
    mul_lanes(E0, a.v[0], a.v[2], a.v[4], a.v[6], b0.v[0]);
    mul_lanes(E1, a.v[0], a.v[2], a.v[4], a.v[6], b1.v[0]);
    mul_lanes(O0, a.v[1], a.v[3], a.v[5], a.v[7], b0.v[0]);
    mul_lanes(O1, a.v[1], a.v[3], a.v[5], a.v[7], b1.v[0]);
    mad_lanes_nc(O0, c.v[1], c.v[3], c.v[5], c.v[7], d0.v[0]);
    mad_lanes_nc(O1, c.v[1], c.v[3], c.v[5], c.v[7], d1.v[0]);
    mad_lanes(E0, O0[7], c.v[0], c.v[2], c.v[4], c.v[6], d0.v[0]);
    mad_lanes(E1, O1[7], c.v[0], c.v[2], c.v[4], c.v[6], d1.v[0]);
    m0 = E0[0] * BN254_INV32;
    m1 = E1[0] * BN254_INV32;
    mad_p_odd(O0, m0);
    mad_p_odd(O1, m1);
    mad_p_even(E0, O0[7], m0);
    mad_p_even(E1, O1[7], m1);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        uint32_t x0 = E0[1], x1 = E1[1];
        uint32_t N0[8] = {E0[2], E0[3], E0[4], E0[5], E0[6], E0[7], 0u, 0u};
        uint32_t N1[8] = {E1[2], E1[3], E1[4], E1[5], E1[6], E1[7], 0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; k++) { E0[k] = O0[k]; E1[k] = O1[k]; }
        mad_lanes_cin(N0, E0[0], x0, a.v[1], a.v[3], a.v[5], a.v[7], b0.v[i]);
        mad_lanes_cin(N1, E1[0], x1, a.v[1], a.v[3], a.v[5], a.v[7], b1.v[i]);
        mad_lanes_nc(N0, c.v[1], c.v[3], c.v[5], c.v[7], d0.v[i]);
        mad_lanes_nc(N1, c.v[1], c.v[3], c.v[5], c.v[7], d1.v[i]);
        mad_lanes(E0, N0[7], a.v[0], a.v[2], a.v[4], a.v[6], b0.v[i]);
        mad_lanes(E1, N1[7], a.v[0], a.v[2], a.v[4], a.v[6], b1.v[i]);
        mad_lanes(E0, N0[7], c.v[0], c.v[2], c.v[4], c.v[6], d0.v[i]);
        mad_lanes(E1, N1[7], c.v[0], c.v[2], c.v[4], c.v[6], d1.v[i]);
        m0 = E0[0] * BN254_INV32;
        m1 = E1[0] * BN254_INV32;
        mad_p_odd(N0, m0);
        mad_p_odd(N1, m1);
        mad_p_even(E0, N0[7], m0);
        mad_p_even(E1, N1[7], m1);
#pragma unroll
        for (int k = 0; k < 8; k++) { O0[k] = N0[k]; O1[k] = N1[k]; }
    }
    fp s0, s1;
    uint32_t h0[8] = {E0[1], E0[2], E0[3], E0[4], E0[5], E0[6], E0[7], 0u};
    uint32_t h1[8] = {E1[1], E1[2], E1[3], E1[4], E1[5], E1[6], E1[7], 0u};
    (void)add_n(s0.v, O0, h0);
    (void)add_n(s1.v, O1, h1);
    fp_reduce_once(s0);
    fp_reduce_once(s1);
    if (!FACTORS_LE_P) {
        fp_reduce_once(s0);
        fp_reduce_once(s1);
    }
    r0 = s0;
    r1 = s1;
}

// The ONE out-of-line instance of the product.  Operands and result travel in registers (by value:
// 16 in, 8 out), so a call costs a handful of moves while every caller stays small: the hot loops
// of the pairing must fit the 32 KB L1.5 instruction cache, and a fully inlined Miller loop is
// ~180 KB of IMAD code (28 % of issue slots lost to instruction-fetch stalls, profiles/r01).
BN_NOINL fp fp_mul_v(fp a, fp b) {
    fp r;
    fp_mul_inl(r, a, b);
    return r;
}
BN_NOINL fp fp_dot2_v(fp a, fp b, fp c, fp d) {
    fp r;
    fp_dot2_inl(r, a, b, c, d);
    return r;
}
BN_INL void fp_mul(fp &r, const fp &a, const fp &b) { r = fp_mul_v(a, b); }
BN_INL void fp_dot2(fp &r, const fp &a, const fp &b, const fp &c, const fp &d) { r = fp_dot2_v(a, b, c, d); }
BN_INL void fp_sqr(fp &r, const fp &a) { fp_mul(r, a, a); }  // fp.rs:409-412

// small multiples
BN_INL void fp_mul3(fp &r, const fp &a) { fp t; fp_dbl(t, a); fp_add(r, t, a); }
BN_INL void fp_mul9(fp &r, const fp &a) {  // fp.rs:414-416 (mul_by_3b, b3 = 9)
    fp t;
    fp_dbl(t, a);
    fp_dbl(t, t);
    fp_dbl(t, t);
    fp_add(r, t, a);
}

}  // namespace bn

namespace bn {

BN_INL const fp &kfp(const uint32_t (&a)[8]) { return *reinterpret_cast<const fp *>(&a[0]); }
BN_INL void fp_set_one(fp &r) { r = kfp(k::ONE); }

BN_INL void fp_mulx(fp &r, const fp &a, const fp &b) { fp_mul(r, a, b); }

// a^e for a fixed 256-bit exponent (plain integer, little-endian 32-bit words).  Same control
// flow for every thread: the exponent is a compile-time table (ff::Field::pow role,
// call sites fp.rs:228,429).
BN_NOINL void fp_pow(fp &r, const fp &a, const uint32_t (&e)[8]) {
    fp acc = kfp(k::ONE), base = a;
#pragma unroll 1
    for (int i = 255; i >= 0; i--) {
        fp_mul(acc, acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) fp_mul(acc, acc, base);
    }
    r = acc;
}
// a^e for one of the fixed exponents, width-4 sliding window over the odd powers a, a^3, ..., a^15:
// ~250 squarings + ~59 products instead of 255 + ~127.  dig[pos] is the (odd) window value whose
// lowest bit sits at bit pos; the tables are generated (tools/gen_constants.py).
BN_NOINL void fp_pow_sw(fp &r, const fp &a, const uint8_t *dig, int top) {
    fp tbl[8], a2, acc;
    tbl[0] = a;
    fp_mul(a2, a, a);
#pragma unroll 1
    for (int i = 1; i < 8; i++) fp_mul(tbl[i], tbl[i - 1], a2);
    acc = tbl[(dig[top] - 1) >> 1];
#pragma unroll 1
    for (int pos = top - 1; pos >= 0; pos--) {
        fp_mul(acc, acc, acc);
        int d = dig[pos];
        if (d) fp_mul(acc, acc, tbl[(d - 1) >> 1]);
    }
    r = acc;
}
BN_INL void fp_inv_fermat(fp &r, const fp &a) { fp_pow_sw(r, a, k::SW_P_MINUS_2, BN254_SW_P_MINUS_2_TOP); }  // 0 -> 0
// fp.rs:428-431 computes the Euler criterion a^((p-1)/2) in {0, 1}; the same predicate (0 counts as a square) is
// evaluated here as the Jacobi symbol (a / p) by the binary algorithm -- shifts, one subtraction and one conditional
// negation per step on 8 limbs, about 1.4 * 254 steps -- which costs a quarter of the 250-squaring exponentiation and
// none of the multiplier pipe.  The Montgomery factor does not matter: R = 2^256 is a square, so (aR / p) = (a / p).
BN_INL bool fp_is_square_euler(const fp &a) {
    fp t;
    fp_pow_sw(t, a, k::SW_P_MINUS_1_DIV_2, BN254_SW_P_MINUS_1_DIV_2_TOP);
    return fp_is_zero(t) || fp_eq(t, kfp(k::ONE));
}
BN_INL uint32_t ctz32(uint32_t x) {  // x != 0
#if BN_PTX
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
BN_INL uint32_t shr_pair(uint32_t lo, uint32_t hi, uint32_t z) {  // bits [z, z + 32) of hi:lo, 0 <= z < 32
#if BN_PTX
    return __funnelshift_r(lo, hi, z);
#else
    return z ? (lo >> z) | (hi << (32 - z)) : lo;
#endif
}
BN_NOINL bool fp_is_square_jacobi(const fp &x) {
    uint32_t a[8], n[8], d[8];
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = x.v[i];
        n[i] = p_limb(i);
        any |= a[i];
    }
    if (!any) return true;
    uint32_t neg = 0;  // the symbol so far is (-1)^neg
    // every pass removes at least one bit of a + n (at most 2 * 256 passes, plus whole-word shifts); the cap only
    // guarantees termination should a caller ever hand in a value that is not a residue below p
#pragma unroll 1
    for (int pass = 0; pass < 2048; pass++) {
        // a != 0.  Strip the factors of two: (2 / n) = -1 iff n = 3, 5 mod 8.  Whole zero words (32 factors: an even count, no
        // sign change) go in a loop of their own -- as a `continue` branch of this loop its register rotation was merged into
        // the common loop tail and every pass paid for it.  z = 0 only before the first subtraction; the shifts take it.
#pragma unroll 1
        while (a[0] == 0) {
#pragma unroll
            for (int i = 0; i < 7; i++) a[i] = a[i + 1];
            a[7] = 0;
        }
        const uint32_t z = ctz32(a[0]);
#pragma unroll
        for (int i = 0; i < 7; i++) a[i] = shr_pair(a[i], a[i + 1], z);
        a[7] >>= z;
        const uint32_t n8 = n[0] & 7u;
        neg ^= (z & 1u) & (uint32_t)((n8 == 3u) | (n8 == 5u));
        // a odd.  d = a - n; when a < n the roles swap (reciprocity: sign flips iff both are 3 mod 4) and a - n becomes n - a.
        // (both differences as borrow chains and a select: cheaper than negating d limb by limb)
        uint32_t e[8];
        uint32_t lt = sub_n(d, a, n);
        (void)sub_n(e, n, a);
        neg ^= ((a[0] & n[0]) >> 1) & lt & 1u;
        uint32_t zero = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            n[i] = lt ? a[i] : n[i];
            uint32_t s = lt ? e[i] : d[i];
            a[i] = s;
            zero |= s;
        }
        if (!zero) break;  // a == n: the gcd, which is 1 for a prime modulus
    }
    return neg == 0;
}
BN_INL uint32_t shl_pair(uint32_t lo, uint32_t hi, uint32_t z) {  // bits [32 - z, 64 - z) of hi:lo, 0 <= z < 32
#if BN_PTX
    return __funnelshift_l(lo, hi, z);
#else
    return z ? (hi << z) | (lo >> (32 - z)) : hi;
#endif
}
// fp.rs:207-210 (the crate calls crypto-bigint's inv_odd_mod): a^-1, 0 -> 0.  Same value as a^(p-2), computed by a
// binary extended GCD in the "almost inverse" form (Kaliski): the loop keeps
//     x A = sigma a 2^k,   x N = -sigma n 2^k   (mod p),    a N + n A = p
// strips the trailing zeros of a (N doubles instead of A halving, so there is no modular halving and several bits
// go at once), then subtracts the smaller of the two odd values from the larger.  About 0.7 * 254 subtraction steps
// of ~70 ALU instructions and no multiplier work; the factor 2^k (and the Montgomery factor) is removed at the end by
// two products.
BN_NOINL void fp_inv_bingcd(fp &r, const fp &x) {
    uint32_t a[8], n[8], A[8], N[8], d[8];
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = x.v[i];
        n[i] = p_limb(i);
        A[i] = 0;
        N[i] = 0;
        any |= a[i];
    }
    if (!any) {
        fp_set_zero(r);
        return;
    }
    A[0] = 1;
    uint32_t sigma_neg = 0, kk = 0;  // sigma = (-1)^sigma_neg
    // every pass removes at least one bit of a + n (at most 2 * 256 passes, plus whole-word shifts); the cap only
    // guarantees termination should a caller ever hand in a value that is not a residue below p
#pragma unroll 1
    for (int pass = 0; pass < 2048; pass++) {
        // 32 trailing zeros at once, in a loop of their own (see fp_is_square_jacobi); z = 0 only before the first subtraction
#pragma unroll 1
        while (a[0] == 0) {
#pragma unroll
            for (int i = 0; i < 7; i++) a[i] = a[i + 1];
            a[7] = 0;
#pragma unroll
            for (int i = 7; i > 0; i--) N[i] = N[i - 1];
            N[0] = 0;
            kk += 32;
        }
        const uint32_t z = ctz32(a[0]);
#pragma unroll
        for (int i = 0; i < 7; i++) a[i] = shr_pair(a[i], a[i + 1], z);
        a[7] >>= z;
#pragma unroll
        for (int i = 7; i > 0; i--) N[i] = shl_pair(N[i - 1], N[i], z);
        N[0] <<= z;
        kk += z;
        // a, n odd.  d = a - n; a < n swaps the roles: (a, A) <- (n - a, A + N), (n, N) <- (a, A), sigma <- -sigma
        uint32_t lt = sub_n(d, a, n);
        uint32_t e[8], S[8];
        (void)sub_n(e, n, a);
        (void)add_n(S, A, N);
        sigma_neg ^= lt & 1u;
        uint32_t zero = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            n[i] = lt ? a[i] : n[i];
            N[i] = lt ? A[i] : N[i];
            A[i] = S[i];
            uint32_t s2 = lt ? e[i] : d[i];
            a[i] = s2;
            zero |= s2;
        }
        if (!zero) break;  // a == n == gcd = 1:  x N = -sigma 2^k
    }
    // y = x^-1 2^k = -sigma N mod p  (N <= p)
    fp y, t;
    uint32_t ge = ~sub_p(y.v, N);  // N - p when N >= p (only N == p is possible; kept general)
#pragma unroll
    for (int i = 0; i < 8; i++) y.v[i] = ge ? y.v[i] : N[i];
    if (!sigma_neg) fp_neg(y, y);
    // x = a_true R: y = a_true^-1 R^-1 2^k.  t = y R = a_true^-1 2^k; the result a_true^-1 R = t 2^(256 - k).
    fp_mul(t, kfp(k::R2), y);
    if (kk > 256) {
        uint32_t j = 512 - kk;  // 4 <= j <= 255: t * 2^j / R
        fp pw;
        fp_set_zero(pw);
        pw.v[j >> 5] = 1u << (j & 31);
        fp_mul(r, t, pw);  // 2^j may exceed p: the row-scanned operand may be unreduced
    } else {
#pragma unroll 1
        for (uint32_t e = kk; e < 256; e++) fp_dbl(t, t);
        r = t;
    }
}
BN_INL void fp_inv(fp &r, const fp &a) {
    fp_inv_bingcd(r, a);
}
BN_INL bool fp_is_square(const fp &a) {
    return fp_is_square_jacobi(a);
}
// candidate root a^((p+1)/4) (fp.rs:212-243 with v = 1); returns whether it squares to a
BN_INL bool fp_sqrt(fp &r, const fp &a) {
    fp y, t;
    fp_pow_sw(y, a, k::SW_P_PLUS_1_DIV_4, BN254_SW_P_PLUS_1_DIV_4_TOP);
    fp_mul(t, y, y);
    r = y;
    return fp_eq(t, a);
}
BN_INL void fp_pow_p_minus_3_div_4(fp &r, const fp &a) { fp_pow_sw(r, a, k::SW_P_MINUS_3_DIV_4, BN254_SW_P_MINUS_3_DIV_4_TOP); }

// Montgomery <-> canonical
// (the multiplicand scanned row-wise -- second argument -- may be any 256-bit value; the first must be < p)
BN_INL void fp_to_mont(fp &r, const fp &a) { fp_mul(r, kfp(k::R2), a); }
BN_INL void fp_from_mont(fp &r, const fp &a) {
    fp one;
    fp_set_zero(one);
    one.v[0] = 1;
    fp_mul(r, a, one);
}
BN_INL uint32_t fp_parity(const fp &a) {  // sgn0, fp.rs:164-168
    fp c;
    fp_from_mont(c, a);
    return c.v[0] & 1u;
}
BN_INL bool fp_raw_lt_p(const fp &a) {
    uint32_t t[8];
    return sub_p(t, a.v) != 0;
}

// 32 big-endian bytes -> canonical limbs (no range check)
BN_INL void fp_raw_from_be(fp &r, const uint8_t *b) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t *q = b + 4 * (7 - i);
        r.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
}
// fp.rs:246-254: rejects values >= p
BN_INL bool fp_from_be(fp &r, const uint8_t *b) {
    fp c;
    fp_raw_from_be(c, b);
    bool ok = fp_raw_lt_p(c);
    fp_to_mont(r, c);
    return ok;
}
// fp.rs:256-260
BN_INL void fp_to_be(uint8_t *b, const fp &a) {
    fp c;
    fp_from_mont(c, a);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint8_t *q = b + 4 * (7 - i);
        q[0] = (uint8_t)(c.v[i] >> 24);
        q[1] = (uint8_t)(c.v[i] >> 16);
        q[2] = (uint8_t)(c.v[i] >> 8);
        q[3] = (uint8_t)c.v[i];
    }
}
// fp.rs:115-122: 48 big-endian bytes (hi 16 || lo 32) mod p, straight into Montgomery form:
// mont(hi*2^256 + lo) = montmul(hi, R^3) + montmul(lo, R^2)
BN_INL void fp_from_okm(fp &r, const uint8_t *b) {
    fp hi, lo, t;
    fp_set_zero(hi);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint8_t *q = b + 4 * (3 - i);
        hi.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    fp_raw_from_be(lo, b + 16);
    fp_mul(hi, kfp(k::R3), hi);
    fp_mul(t, kfp(k::R2), lo);  // lo may exceed p: fine as the row-scanned operand, result is reduced
    fp_add(r, hi, t);
}

}  // namespace bn
