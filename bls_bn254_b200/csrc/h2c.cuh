// h2c.cuh -- RFC 9380 hash-to-curve for BN254: expand_message_xmd (SHA-256), hash_to_field,
// the Shallue-van de Woestijne maps to G1 and G2, and the hash / encode compositions.
//
// Role of /root/reference/src/inner_types/fp.rs:284-371 (SvdW -> G1), 433-458 (Fp::hash/encode),
// fp2.rs:221-287 (SvdW -> G2), 463-488 (Fp2::hash/encode), g1.rs:910-928, g2.rs:919-936.
// expand_message_xmd lives in the un-vendored dependency elliptic-curve 0.13.8
// (hash2curve::ExpandMsgXmd<Sha256>); RFC 9380 section 5.3.1 is restated here and pinned by the
// `u` / `u0` / `u1` fields of the reference's known-answer tests (g1.rs:999-1119).
#pragma once
#include "curve.cuh"

namespace bn {

// ------------------------------------------------------------------------------------- SHA-256
struct sha256 {
    uint32_t h[8];
    uint32_t w[16];  // current block, big-endian words
    uint32_t fill;   // bytes in the current block
    uint64_t len;    // total bytes
};

BN_INL uint32_t rotr(uint32_t x, int n) {
#if BN_PTX
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

BN_NOINL void sha256_compress(uint32_t h[8], uint32_t w[16]) {
    const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + wi;
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

BN_INL void sha256_init(sha256 &s) {
    s.h[0] = 0x6a09e667; s.h[1] = 0xbb67ae85; s.h[2] = 0x3c6ef372; s.h[3] = 0xa54ff53a;
    s.h[4] = 0x510e527f; s.h[5] = 0x9b05688c; s.h[6] = 0x1f83d9ab; s.h[7] = 0x5be0cd19;
#pragma unroll
    for (int i = 0; i < 16; i++) s.w[i] = 0;
    s.fill = 0;
    s.len = 0;
}
BN_INL void sha256_put(sha256 &s, uint8_t byte) {
    uint32_t idx = s.fill >> 2, sh = 24 - 8 * (s.fill & 3);
    s.w[idx] |= (uint32_t)byte << sh;
    s.fill++;
    s.len++;
    if (s.fill == 64) {
        sha256_compress(s.h, s.w);
#pragma unroll
        for (int i = 0; i < 16; i++) s.w[i] = 0;
        s.fill = 0;
    }
}
// One big-endian word at a word boundary of the block (fill % 4 == 0): a single store instead of four read-modify-writes.
BN_INL void sha256_put_word(sha256 &s, uint32_t word) {
    s.w[s.fill >> 2] = word;
    s.fill += 4;
    s.len += 4;
    if (s.fill == 64) {
        sha256_compress(s.h, s.w);
#pragma unroll
        for (int i = 0; i < 16; i++) s.w[i] = 0;
        s.fill = 0;
    }
}
BN_INL uint32_t load_be32(const uint8_t *d) {
    if ((reinterpret_cast<uintptr_t>(d) & 3u) == 0) {
        const uint32_t v = *reinterpret_cast<const uint32_t *>(d);
#if BN_PTX
        return __byte_perm(v, 0, 0x0123);
#else
        return __builtin_bswap32(v);
#endif
    }
    return (uint32_t)d[0] << 24 | (uint32_t)d[1] << 16 | (uint32_t)d[2] << 8 | (uint32_t)d[3];
}
// word-wise absorption: bytes up to the next word boundary of the block, whole words, the remaining bytes
BN_INL void sha256_update(sha256 &s, const uint8_t *d, size_t n) {
#pragma unroll 1
    while (n && (s.fill & 3u)) {
        sha256_put(s, *d++);
        n--;
    }
#pragma unroll 1
    while (n >= 4) {
        sha256_put_word(s, load_be32(d));
        d += 4;
        n -= 4;
    }
#pragma unroll 1
    while (n) {
        sha256_put(s, *d++);
        n--;
    }
}
// padding and the last compression(s); the digest stays in s.h as eight big-endian words
BN_INL void sha256_finish(sha256 &s) {
    uint64_t bits = s.len * 8;
    sha256_put(s, 0x80);
#pragma unroll 1
    while (s.fill != 56) {
        if (s.fill & 3u) sha256_put(s, 0);
        else sha256_put_word(s, 0);
    }
    s.w[14] = (uint32_t)(bits >> 32);
    s.w[15] = (uint32_t)bits;
    sha256_compress(s.h, s.w);
}
BN_INL void sha256_final(sha256 &s, uint8_t out[32]) {
    sha256_finish(s);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(s.h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s.h[i] >> 8);
        out[4 * i + 3] = (uint8_t)s.h[i];
    }
}
// SHA-256 state after the 64 zero bytes of Z_pad (RFC 9380 5.3.1 step 7): the first block of every b_0 is constant, so its
// compression is done once, here (tools: the value is the compression function applied to the IV and a zero block).
BN_INL void sha256_init_zpad(sha256 &s) {
    s.h[0] = 0xda5698be; s.h[1] = 0x17b9b469; s.h[2] = 0x62335799; s.h[3] = 0x779fbeca;
    s.h[4] = 0x8ce5d491; s.h[5] = 0xc0d26243; s.h[6] = 0xbafef9ea; s.h[7] = 0x1837a9d8;
#pragma unroll
    for (int i = 0; i < 16; i++) s.w[i] = 0;
    s.fill = 0;
    s.len = 64;
}

// RFC 9380 5.3.1 with H = SHA-256 (b = 32, s = 64); n <= 255*32.  dst_len <= 255 (an oversize DST
// is replaced by H("H2C-OVERSIZE-DST-" || DST) before this is called, see oversize_dst()).
// b_0 and the b_i stay as big-endian WORDS: b_0 xor b_(i-1) is fed to the next hash eight words at a time.
BN_INL void expand_message_xmd(uint8_t *out, uint32_t n, const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len) {
    uint32_t b0[8], bi[8];
    sha256 s;
    uint32_t ell = (n + 31) / 32;
    sha256_init_zpad(s);
    sha256_update(s, msg, msg_len);
    sha256_put(s, (uint8_t)(n >> 8));
    sha256_put(s, (uint8_t)n);
    sha256_put(s, 0);
    sha256_update(s, dst, dst_len);
    sha256_put(s, (uint8_t)dst_len);
    sha256_finish(s);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        b0[j] = s.h[j];
        bi[j] = 0;
    }
#pragma unroll 1
    for (uint32_t i = 1; i <= ell; i++) {
        sha256_init(s);
#pragma unroll
        for (int j = 0; j < 8; j++) sha256_put_word(s, b0[j] ^ bi[j]);
        sha256_put(s, (uint8_t)i);
        sha256_update(s, dst, dst_len);
        sha256_put(s, (uint8_t)dst_len);
        sha256_finish(s);
#pragma unroll
        for (int j = 0; j < 8; j++) bi[j] = s.h[j];
        uint32_t off = (i - 1) * 32, take = n - off < 32 ? n - off : 32;
#pragma unroll 1
        for (uint32_t j = 0; j < take; j++) out[off + j] = (uint8_t)(bi[j >> 2] >> (24 - 8 * (j & 3u)));
    }
}
BN_INL void oversize_dst(uint8_t out[32], const uint8_t *dst, size_t dst_len) {
    const char tag[18] = "H2C-OVERSIZE-DST-";
    sha256 s;
    sha256_init(s);
#pragma unroll 1
    for (int i = 0; i < 17; i++) sha256_put(s, (uint8_t)tag[i]);
    sha256_update(s, dst, dst_len);
    sha256_final(s, out);
}

// --------------------------------------------------------------------------------- SvdW maps
// fp.rs:284-371 (Z = 1, c1 = g(Z) = 4), split in two so that the inversions of both maps of a
// hash (and only those) share ONE field inversion (Montgomery's trick, 3 extra products):
//   svdw_g1_pre : tv1 = 1 - 4u^2, tv2 = 1 + 4u^2, tv3 = tv1 tv2     (the value to invert)
//   svdw_g1_post: everything after tv3 = inv0(tv1 tv2)
// Returns false where the reference panics (tv1*tv2 == 0, i.e. u = +-1/2; fp.rs:325).
BN_INL void svdw_g1_pre(fp &tv1, fp &tv2, fp &tv3, const fp &u) {
    const fp &one = kfp(k::ONE);
    fp_mulx(tv1, u, u);
    fp_dbl(tv1, tv1);
    fp_dbl(tv1, tv1);  // * c1 = 4
    fp_add(tv2, one, tv1);
    fp_sub(tv1, one, tv1);
    fp_mulx(tv3, tv1, tv2);
}
#ifndef BN_SVDW1
#define BN_SVDW1 BN_NOINL
#endif
BN_SVDW1 bool svdw_g1_post(g1a &out, const fp &u, const fp &tv1, const fp &tv2, const fp &tv3 /* inverted */) {
    fp tv4, x1, x2, x3, gx1, gx2, gx, x, y;
    const fp &one = kfp(k::ONE);
    bool ok = !fp_is_zero(tv3);
    fp_mulx(tv4, u, tv1);
    fp_mulx(tv4, tv4, tv3);
    fp_mulx(tv4, tv4, kfp(k::SVDW1_C3));
    fp_sub(x1, kfp(k::SVDW1_C2), tv4);
    fp_mulx(gx1, x1, x1); fp_mulx(gx1, gx1, x1); fp_add(gx1, gx1, kfp(k::THREE));
    fp_add(x2, kfp(k::SVDW1_C2), tv4);
    fp_mulx(gx2, x2, x2); fp_mulx(gx2, gx2, x2); fp_add(gx2, gx2, kfp(k::THREE));
    fp_mulx(x3, tv2, tv2); fp_mulx(x3, x3, tv3); fp_mulx(x3, x3, x3); fp_mulx(x3, x3, kfp(k::SVDW1_C4));
    fp_add(x3, x3, one);
    bool e1 = fp_is_square(gx1);
    bool e2 = fp_is_square(gx2) & !e1;
    fp_select(x, e1, x1, x3);
    fp_select(x, e2, x2, x);
    fp_mulx(gx, x, x); fp_mulx(gx, gx, x); fp_add(gx, gx, kfp(k::THREE));
    ok &= fp_sqrt(y, gx);
    fp ny;
    fp_neg(ny, y);
    fp_select(y, (fp_parity(u) ^ fp_parity(y)) != 0, ny, y);
    out.x = x;
    out.y = y;
    out.inf = false;
    return ok;
}
BN_INL bool svdw_g1(g1a &out, const fp &u) {
    fp tv1, tv2, tv3;
    svdw_g1_pre(tv1, tv2, tv3, u);
    fp_inv(tv3, tv3);
    return svdw_g1_post(out, u, tv1, tv2, tv3);
}
// fp2.rs:221-287, same split
BN_INL void svdw_g2_pre(fp2 &tv1, fp2 &tv2, fp2 &tv3, const fp2 &u) {
    fp2 one;
    fp2_set_one(one);
    fp2_sqr(tv1, u);
    fp2_mul(tv1, tv1, kfp2(k::SVDW2_C1));
    fp2_add(tv2, one, tv1);
    fp2_sub(tv1, one, tv1);
    fp2_mul(tv3, tv1, tv2);
}
// Inline on purpose: as an out-of-line function with five reference parameters this body produced wrong
// points on sm_100a (nvcc 12.9) while the identical source is correct on the host and when inlined
// (bisected on the GPU in round 1, see profiles/r01_notes.md); the parity suite guards it.
#ifndef BN_SVDW
#define BN_SVDW BN_INL
#endif
BN_SVDW bool svdw_g2_post(g2a &out, const fp2 &u, const fp2 &tv1, const fp2 &tv2, const fp2 &tv3 /* inverted */) {
    fp2 tv4, x1, x2, x3, gx1, gx2, gx, x, y, one;
    fp2_set_one(one);
    bool ok = !fp2_is_zero(tv3);
    fp2_mul(tv4, u, tv1);
    fp2_mul(tv4, tv4, tv3);
    fp2_mul(tv4, tv4, kfp2(k::SVDW2_C3));
    fp2_sub(x1, kfp2(k::SVDW2_C2), tv4);
    fp2_sqr(gx1, x1); fp2_mul(gx1, gx1, x1); fp2_add(gx1, gx1, kfp2(k::B2));
    fp2_add(x2, kfp2(k::SVDW2_C2), tv4);
    fp2_sqr(gx2, x2); fp2_mul(gx2, gx2, x2); fp2_add(gx2, gx2, kfp2(k::B2));
    fp2_sqr(x3, tv2); fp2_mul(x3, x3, tv3); fp2_sqr(x3, x3); fp2_mul(x3, x3, kfp2(k::SVDW2_C4));
    fp2_add(x3, x3, one);
    bool e1 = fp2_is_square(gx1);
    bool e2 = fp2_is_square(gx2) & !e1;
    fp2_select(x, e1, x1, x3);
    fp2_select(x, e2, x2, x);
    fp2_sqr(gx, x); fp2_mul(gx, gx, x); fp2_add(gx, gx, kfp2(k::B2));
    ok &= fp2_sqrt(y, gx);
    fp2 ny;
    fp2_neg(ny, y);
    fp2_select(y, (fp2_sgn0(u) ^ fp2_sgn0(y)) != 0, ny, y);
    out.x = x;
    out.y = y;
    out.inf = false;
    return ok;
}
BN_INL bool svdw_g2(g2a &out, const fp2 &u) {
    fp2 tv1, tv2, tv3;
    svdw_g2_pre(tv1, tv2, tv3, u);
    fp2_inv(tv3, tv3);
    return svdw_g2_post(out, u, tv1, tv2, tv3);
}

// hash (count = 2, g1.rs:910-919) / encode (count = 1, g1.rs:922-928).  False on the SvdW exception.
BN_INL bool hash_to_g1(g1a &out, const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len, int count) {
    uint8_t okm[96];
    fp u;
    g1a q0, q1;
    expand_message_xmd(okm, 48u * (uint32_t)count, msg, msg_len, dst, dst_len);
    fp_from_okm(u, okm);
    if (count == 1) {
        bool ok1 = svdw_g1(q0, u);
        out = q0;
        return ok1;
    }
    // both maps share one inversion: 1/a = b (ab)^-1, 1/b = a (ab)^-1 (a zero makes both "inverses"
    // zero, and the item is reported as the SvdW exception, as for a single zero)
    fp u1, a1, a2, a3, b1, b2, b3, t;
    fp_from_okm(u1, okm + 48);
    svdw_g1_pre(a1, a2, a3, u);
    svdw_g1_pre(b1, b2, b3, u1);
    fp_mulx(t, a3, b3);
    fp_inv(t, t);
    fp ia, ib;
    fp_mulx(ia, t, b3);
    fp_mulx(ib, t, a3);
    bool ok = svdw_g1_post(q0, u, a1, a2, ia);
    ok &= svdw_g1_post(q1, u1, b1, b2, ib);
    g1p a, b;
    pt_from_affine(a, q0.x, q0.y, false);
    pt_from_affine(b, q1.x, q1.y, false);
    pt_add(a, a, b);
    pt_to_affine(out.x, out.y, out.inf, a);
    return ok;
}
// g2.rs:919-936 (c0 from the first 48 bytes, c1 from the next 48: fp2.rs:454-461)
BN_INL bool hash_to_g2(g2a &out, const uint8_t *msg, size_t msg_len, const uint8_t *dst, uint32_t dst_len, int count) {
    uint8_t okm[192];
    fp2 u;
    g2a q;
    g2p a, b;
    expand_message_xmd(okm, 96u * (uint32_t)count, msg, msg_len, dst, dst_len);
    fp_from_okm(u.c0, okm);
    fp_from_okm(u.c1, okm + 48);
    bool ok;
    if (count == 2) {
        fp2 u1, a1, a2, a3, b1, b2, b3, t, ia, ib;
        fp_from_okm(u1.c0, okm + 96);
        fp_from_okm(u1.c1, okm + 144);
        svdw_g2_pre(a1, a2, a3, u);
        svdw_g2_pre(b1, b2, b3, u1);
        fp2_mul(t, a3, b3);
        fp2_inv(t, t);
        fp2_mul(ia, t, b3);
        fp2_mul(ib, t, a3);
        ok = svdw_g2_post(q, u, a1, a2, ia);
        pt_from_affine(a, q.x, q.y, false);
        ok &= svdw_g2_post(q, u1, b1, b2, ib);
        pt_from_affine(b, q.x, q.y, false);
        pt_add(a, a, b);
    } else {
        ok = svdw_g2(q, u);
        pt_from_affine(a, q.x, q.y, false);
    }
    g2_clear_cofactor(a, a);
    pt_to_affine(out.x, out.y, out.inf, a);
    return ok;
}

}  // namespace bn
