// vm.cuh -- the pairing "slot machine": Miller loops and the final exponentiation as PROGRAMS over shared-memory slots.
//
// Role of /root/reference/src/inner_types/pairings.rs:50-178 (final exponentiation), 760-962 (pairing, multi_miller_loop,
// ell, doubling_step, addition_step), fp6.rs:123-287, fp12.rs:120-219 -- the same functions as pairing.cuh / tower.cuh, bit for
// bit (every value is a canonical Montgomery residue after every operation), in the shape the B200 wants:
//
//   * Round 1's pairing kernels were call sequences of Fp2 leaves on per-thread LOCAL memory: 1.6-2.7 KB of stack per
//     thread x 896 threads per SM does not fit L1 (hit rate 50-65 %, 29 GB of DRAM write-back per Miller launch) and the
//     multiplier pipe idled at 66 % (Miller) / 54 % (final exponentiation) behind `long_scoreboard` stalls.
//   * Here a lane pair owns one pairing (even lane: the c0 half of every Fp2 value, odd lane: c1, exchanged by SHFL as in
//     tower.cuh: fp2l) and EVERY live value sits in a shared-memory SLOT: 32 bytes per lane at [slot][half][lane], so each
//     LDS.128 / STS.128 of a warp is conflict free and nothing long-lived ever leaves the SM.  ~20 slots per lane at one
//     warp per CTA is 20 KB per CTA: 11 warps per SM stay resident, enough to keep the multiplier pipe busy once no warp
//     waits on L2.
//   * The tower algebra is DATA (vm_prog.cuh, generated and checked by tools/gen_vm_programs.py): a 16-byte instruction is a
//     whole fused step  R = 2^k (xi^e ((sum X_i)(sum Y_j) + sum Q_k) + sum R_l), so Karatsuba's pre-additions and
//     post-subtractions run in the shadow of the product instead of being separate load / add / store round trips, and the
//     kernel's code is this interpreter (a few KB, resident in the instruction cache) instead of 50 KB of call sequences.
//
// The interpreter is generic over the slot memory (shared memory on the device, a plain array in tests/hostsim).
#pragma once
#include "pairing.cuh"
#include "vm_prog.cuh"

namespace bn {
namespace vm {

#ifndef BN_VM_TPB
#define BN_VM_TPB 32   // threads per CTA of the slot-machine kernels: one warp = 16 pairings
#endif

// what a routine can reach besides its slots: the item's Fp scalars (G1 coordinates) and Fp2 inputs (G2 coordinates) in
// HBM (Montgomery records of items.cuh), the per-item scratch of the final exponentiation, and two driver registers
struct Env {
    const fp *p0, *p1;    // scalars 0, 1 = p0[0], p0[1] (P0.x, P0.y); scalars 2, 3 = p1[0], p1[1] (P1.x, P1.y)
    const fp2 *q0, *q1;   // inputs 0, 1 = q0[0], q0[1] (Q0.x, Q0.y); inputs 2, 3 = q1[0], q1[1]
    fp2 *scratch;         // entry e of this item lives at scratch[e * scratch_stride]
    size_t scratch_stride;
    int line;             // line counter of the fixed-point table ("rel" constants: index + 3 * line)
    int var;              // scratch variable of FE_LDF / FE_LDG / FE_STF ("rel" scratch entries: index + 6 * var)
};

#if defined(__CUDACC__)
// This lane's column of the CTA's slot array as a 32-bit SHARED-space address: slot s = 16 bytes at base + 2 s TPB 16 and
// 16 bytes at base + (2 s + 1) TPB 16.  Explicit ld.shared / st.shared: the interpreter is an out-of-line function, and a
// pointer parameter would degrade to generic LD.E / ST.E (what sank round 1's -DBN_SMEM_ACC experiment).
struct Slots {
    uint32_t base;
    __device__ __forceinline__ fp ld(int s) const {
        fp r;
        const uint32_t a = base + (uint32_t)s * (32u * BN_VM_TPB);
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                     : "r"(a + 16u * BN_VM_TPB)
                     : "memory");
        return r;
    }
    __device__ __forceinline__ void st(int s, const fp &v) const {
        const uint32_t a = base + (uint32_t)s * (32u * BN_VM_TPB);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.v[0]), "r"(v.v[1]), "r"(v.v[2]), "r"(v.v[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16u * BN_VM_TPB), "r"(v.v[4]), "r"(v.v[5]), "r"(v.v[6]), "r"(v.v[7])
                     : "memory");
    }
};
__device__ __forceinline__ Slots make_slots(const void *smem_base) {
    Slots m;
    m.base = (uint32_t)__cvta_generic_to_shared(smem_base) + 16u * threadIdx.x;
    return m;
}
#else
struct Slots {            // host build: one array per emulated lane
    fp *a;
    fp ld(int s) const { return a[s]; }
    void st(int s, const fp &v) const { a[s] = v; }
};
#endif

// (9 + u) v for this lane's half: even lane 9 v0 - v1, odd lane 9 v1 + v0
BN_NOINL fp lane_xi_v(fp v, bool odd) {
    fp pv, n, t, u, r;
    lane_xchg(pv, v);
    fp_neg(n, pv);
    fp_mul9(t, v);
    fp_select(u, odd, pv, n);
    fp_add(r, t, u);
    return r;
}
BN_INL void lane_xi(fp &r, const fp &v, bool odd) { r = lane_xi_v(v, odd); }

// the descriptor queue of an instruction: 12 bytes in three words, consumed from the low end
struct Desc {
    uint32_t w1, w2, w3;
    BN_INL uint32_t pop() {
        const uint32_t d = w1 & 0xffu;
        w1 = (w1 >> 8) | (w2 << 24);
        w2 = (w2 >> 8) | (w3 << 24);
        w3 >>= 8;
        return d;
    }
};

// sum of the next n operand terms (n >= 1): +-[xi] slot
BN_INL fp gather(const Slots M, Desc &D, int n, bool odd) {
    fp acc;
    fp_set_zero(acc);
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const uint32_t d = D.pop();
        fp v = M.ld((int)(d & 63u));
        if (d & 0x40u) lane_xi(v, v, odd);
        if (i == 0) {
            if (d & 0x80u) fp_neg(acc, v);
            else acc = v;
        } else {
            if (d & 0x80u) fp_sub(acc, acc, v);
            else fp_add(acc, acc, v);
        }
    }
    return acc;
}

// this lane's half of an Fp2 constant / input (address select, one load path)
BN_INL fp const_half(int idx, bool odd) {
    const uint32_t *p = &VM_CONST[idx][odd ? 1 : 0][0];
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = p[i];
    return r;
}
BN_INL fp global_half(const fp2 *q, bool odd) {
    const fp *p = odd ? &q->c1 : &q->c0;
    return *p;
}

// Runs one routine.  All lanes of a warp execute the same instruction stream (the program is warp uniform), so the lane
// exchanges inside are full-mask shuffles.
BN_NOINL void run(const Slots M, const Env E, int routine) {
    const bool odd = lane_par() != 0;
    int pc = VM_DIR[routine][0];
    const int end = pc + VM_DIR[routine][1];
#pragma unroll 1
    for (; pc < end; pc++) {
        const uint32_t w0 = VM_PROG[pc][0];
        Desc D = {VM_PROG[pc][1], VM_PROG[pc][2], VM_PROG[pc][3]};
        const int op = (int)(w0 & 15u), d = (int)((w0 >> 4) & 63u);
        const int nx = (int)((w0 >> 10) & 7u), ny = (int)((w0 >> 13) & 7u), nq = (int)((w0 >> 16) & 3u), nr = (int)((w0 >> 18) & 7u);
        const bool xi = (w0 >> 21) & 1u, conjx = (w0 >> 24) & 1u, rel = (w0 >> 25) & 1u;
        const int dbl = (int)((w0 >> 22) & 3u), aux = (int)(w0 >> 26);
        const int cidx = (int)(D.w3 >> 16) + (rel ? 3 * E.line : 0);
        if (op <= OP_MULFPC) {
            // ---- product class: p = X * (Y | X | scalar | constant) ------------------------------------------------
            fp x, y, p;
            if (op == OP_MULFPC) x = const_half(cidx, odd);
            else x = gather(M, D, nx, odd);
            if (conjx) {
                fp n;
                fp_neg(n, x);
                fp_select(x, odd, n, x);
            }
            if (op == OP_MUL || op == OP_MULC) {
                // even lane: x0 y0 + x1 (-y1);  odd lane: x1 y0 + x0 y1  -- one fused dot product per lane
                fp px, py, npy, u, w;
                lane_xchg(px, x);
                if (op == OP_MUL) {
                    y = gather(M, D, ny, odd);
                    lane_xchg(py, y);
                } else {
                    y = const_half(cidx, odd);
                    py = const_half(cidx, !odd);
                }
                fp_neg(npy, py);
                fp_select(u, odd, py, y);
                fp_select(w, odd, y, npy);
                fp_dot2_inl(p, x, u, px, w);
            } else {
                if (op == OP_SQR) {
                    // even lane: (x0 + x1)(x0 - x1);  odd lane: (x0 + x0) x1
                    fp px, s, dd;
                    lane_xchg(px, x);
                    fp_select(s, odd, px, x);
                    fp_sub(dd, x, px);
                    fp_select(y, odd, x, dd);
                    fp_add(x, s, px);
                } else {
                    y = ((aux & 2) ? E.p1 : E.p0)[aux & 1];   // OP_MULFP / OP_MULFPC: an Fp scalar of the item, the same for both lanes
                }
                fp_mul_inl(p, x, y);
            }
            if (nq) {
                fp q = gather(M, D, nq, odd);
                fp_add(p, p, q);
            }
            if (xi) lane_xi(p, p, odd);
            if (nr) {
                fp r = gather(M, D, nr, odd);
                fp_add(p, p, r);
            }
#pragma unroll 1
            for (int k = 0; k < dbl; k++) fp_dbl(p, p);
            M.st(d, p);
        } else if (op == OP_LIN) {
            // ---- linear step: sum of up to 8 terms with coefficients 1..4, then 2^dbl, then an optional conjugation ------
            uint32_t coefs = D.w3;
            fp acc;
            fp_set_zero(acc);
#pragma unroll 1
            for (int i = 0; i < nx; i++) {
                const uint32_t de = D.pop();
                fp v = M.ld((int)(de & 63u));
                if (de & 0x40u) lane_xi(v, v, odd);
                const int c = (int)(coefs & 3u);      // coefficient - 1
                coefs >>= 2;
                fp t = v;
                if (c >= 1) fp_dbl(t, v);
                if (c == 2) fp_add(t, t, v);
                if (c == 3) fp_dbl(t, t);
                if (de & 0x80u) fp_sub(acc, acc, t);
                else fp_add(acc, acc, t);
            }
#pragma unroll 1
            for (int k = 0; k < dbl; k++) fp_dbl(acc, acc);
            if (conjx) {
                fp n;
                fp_neg(n, acc);
                fp_select(acc, odd, n, acc);
            }
            M.st(d, acc);
        } else if (op == OP_LDC) {
            M.st(d, const_half(cidx, odd));
        } else if (op == OP_LDG) {
            M.st(d, global_half(((aux & 2) ? E.q1 : E.q0) + (aux & 1), odd));
        } else if (op == OP_INV) {
            fp2l a, r;
            a.v = gather(M, D, nx, odd);
            fp2_inv(r, a);
            M.st(d, r.v);
        } else if (op == OP_FILL) {
            const fp2 *q = E.scratch + (size_t)(aux + (rel ? 6 * E.var : 0)) * E.scratch_stride;
            M.st(d, global_half(q, odd));
        } else if (op == OP_SPILL) {
            fp2 *q = E.scratch + (size_t)(aux + (rel ? 6 * E.var : 0)) * E.scratch_stride;
            fp *p = odd ? &q->c1 : &q->c0;
            *p = M.ld(d);
        }
    }
}

// ---- the WIDE machine: one item per warp, 16 units (lane pairs), bundles of independent instructions ------------------------
// Programs: vmw_prog.cuh (tools/gen_vm_programs.py, "the WIDE machine").  Every unit of the warp takes instruction `unit` of
// the bundle (units beyond the bundle's size idle), reads its operands from the item's slots, the warp synchronises, the
// unit computes and stores, the warp synchronises.  Units run different opcodes, so the warp diverges inside a bundle and
// the pair exchanges name only the two lanes of the unit.  A slot is 64 bytes: [c0 | c1].  Slot ids: 0..11 the Fp12
// registers / twist points as in the sequential machine, W_SC + i the item's Fp scalars (c0 half), W_IN + i its Fp2 inputs,
// W_VAR + 6 v + i the scratch variables of the final exponentiation, W_TMP.. temporaries, W_CONST + c = constant c (read
// from constant memory, never written).
#if defined(__CUDACC__)
#define BN_GDECL __device__ __align__(16) const
#else
#define BN_GDECL alignas(16) static const
#endif
}  // namespace vm
}  // namespace bn
#include "vmw_prog.cuh"
namespace bn {
namespace vm {

#if defined(__CUDACC__)
// The CTA's shared memory: [W_DIR | W_BOFF | the instructions of the routines the kernel runs | one slot array per warp].
// Programs are staged there once per CTA (wide_stage): fetched from global memory, the three dependent loads per bundle
// (bundle offsets, instruction, operands) cost ~3000 cycles -- more than the arithmetic.
constexpr int W_DIR_BYTES = ((R_COUNT * 4 + 15) / 16) * 16;
constexpr int W_BOFF_BYTES = (((int)(sizeof(W_BOFF)) + 15) / 16) * 16;
struct WSlots {
    uint32_t base;      // shared-space address of the item's slot array
    uint32_t prog;      // shared-space address instruction i lives at prog + 32 i (the staged range starts at its first instruction)
    uint32_t tabs;      // shared-space address of the directory copy; the bundle offsets follow it
    __device__ __forceinline__ fp ldh(int s, int half) const {
        fp r;
        const uint32_t a = base + (uint32_t)s * 64u + (uint32_t)half * 32u;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "r"(a + 16u) : "memory");
        return r;
    }
    __device__ __forceinline__ void sth(int s, int half, const fp &v) const {
        const uint32_t a = base + (uint32_t)s * 64u + (uint32_t)half * 32u;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.v[0]), "r"(v.v[1]), "r"(v.v[2]), "r"(v.v[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16u), "r"(v.v[4]), "r"(v.v[5]), "r"(v.v[6]), "r"(v.v[7]) : "memory");
    }
    __device__ __forceinline__ int dir(int routine, int k) const {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(tabs + 4u * (uint32_t)routine + 2u * (uint32_t)k));
        return (int)v;
    }
    __device__ __forceinline__ int boff(int b) const {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(tabs + (uint32_t)W_DIR_BYTES + 2u * (uint32_t)b));
        return (int)v;
    }
    __device__ __forceinline__ void ins(int i, uint32_t (&w)[8]) const {
        const uint32_t a = prog + 32u * (uint32_t)i;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(a + 16u));
    }
};
// bytes of shared memory a CTA of `warps` items needs to run routines [r_lo, r_hi]
__host__ __device__ inline size_t wide_smem_bytes(int r_lo, int r_hi, int warps, int i_lo, int i_hi) {
    (void)r_lo; (void)r_hi;
    return (size_t)W_DIR_BYTES + W_BOFF_BYTES + (size_t)(i_hi - i_lo) * 32 + (size_t)warps * W_NSLOT * 64;
}
// All threads of the CTA: copy the tables and the instructions of routines [r_lo, r_hi] into shared memory, barrier, and
// return this warp's machine.
__device__ __forceinline__ WSlots wide_stage(void *smem, int r_lo, int r_hi) {
    const int i_lo = W_BOFF[W_DIR[r_lo][0]], i_hi = W_BOFF[W_DIR[r_hi][0] + W_DIR[r_hi][1]];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    for (int k = threadIdx.x; k < R_COUNT * 2; k += blockDim.x) tab[k] = W_DIR[k >> 1][k & 1];
    uint16_t *bo = tab + W_DIR_BYTES / 2;
    for (int k = threadIdx.x; k < (int)(sizeof(W_BOFF) / 2); k += blockDim.x) bo[k] = W_BOFF[k];
    uint4 *pr = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(smem) + W_DIR_BYTES + W_BOFF_BYTES);
    const uint4 *src = reinterpret_cast<const uint4 *>(&W_PROG[i_lo][0]);
    for (int k = threadIdx.x; k < 2 * (i_hi - i_lo); k += blockDim.x) pr[k] = src[k];
    __syncthreads();
    WSlots m;
    m.tabs = (uint32_t)__cvta_generic_to_shared(smem);
    m.prog = m.tabs + (uint32_t)(W_DIR_BYTES + W_BOFF_BYTES) - 32u * (uint32_t)i_lo;
    m.base = m.tabs + (uint32_t)(W_DIR_BYTES + W_BOFF_BYTES) + 32u * (uint32_t)(i_hi - i_lo) + (threadIdx.x >> 5) * (uint32_t)(W_NSLOT * 64);
    return m;
}
__device__ __forceinline__ int wide_lane() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ void wide_sync() { __syncwarp(); }
__device__ __forceinline__ void wide_xchg(fp &out, const fp &in) {   // the partner lane's `in`; only the two lanes of the unit take part
    const uint32_t m = 3u << (threadIdx.x & 30u);
#pragma unroll
    for (int i = 0; i < 8; i++) out.v[i] = __shfl_xor_sync(m, in.v[i], 1);
}
#else
extern "C" int bn_host_lane_id();
extern "C" void bn_host_warp_sync();
struct WSlots {            // host build: one array shared by the 32 emulated lanes
    fp (*a)[2];
    fp ldh(int s, int half) const { return a[s][half]; }
    void sth(int s, int half, const fp &v) const { a[s][half] = v; }
    int dir(int routine, int k) const { return W_DIR[routine][k]; }
    int boff(int b) const { return W_BOFF[b]; }
    void ins(int i, uint32_t (&w)[8]) const {
        for (int k = 0; k < 8; k++) w[k] = W_PROG[i][k];
    }
};
inline int wide_lane() { return bn_host_lane_id(); }
inline void wide_sync() { bn_host_warp_sync(); }
inline void wide_xchg(fp &out, const fp &in) { bn_host_lane_xchg(out.v, in.v, 8); }
#endif

BN_INL fp wide_xi(const fp &v, bool odd) {     // (9 + u) v for this lane's half
    fp pv, n, t, u, r;
    wide_xchg(pv, v);
    fp_neg(n, pv);
    fp_mul9(t, v);
    fp_select(u, odd, pv, n);
    fp_add(r, t, u);
    return r;
}
// descriptor queue of a wide instruction: twelve 16-bit descriptors in six words, consumed from the low end
struct WDesc {
    uint32_t w[6];
    BN_INL uint32_t pop() {
        const uint32_t d = w[0] & 0xffffu;
#pragma unroll
        for (int i = 0; i < 5; i++) w[i] = (w[i] >> 16) | (w[i + 1] << 16);
        w[5] >>= 16;
        return d;
    }
};
// ---- lazy linear combinations.  A gather sums up to twelve terms +- [xi] c * slot (c <= 4).  Reducing after every term costs
// ~30 dependent instructions per term (and a multiplication by xi = 9 + u a lane exchange and four modular operations); here
// the terms are accumulated UNREDUCED in a signed 288-bit accumulator -- a term is one multiply-add by a small constant, and
// xi * v = (9 v0 - v1) + (9 v1 + v0) u reads both halves of v from the slot instead of exchanging -- and one reduction ends
// the gather: add 512 p (every accumulator stays within +- 480 p), estimate the quotient from the top 64 bits, subtract q p,
// at most two conditional subtractions.
struct WAcc {
    uint32_t a[9];
};
BN_INL void wacc_zero(WAcc &A) {
#pragma unroll
    for (int i = 0; i < 9; i++) A.a[i] = 0;
}
// A += (neg ? -1 : 1) * k * v   (two's complement; branch free)
BN_INL void wacc_mad(WAcc &A, const fp &v, uint32_t k, bool neg) {
    uint32_t w[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t t = (uint64_t)v.v[i] * k + c;
        w[i] = (uint32_t)t;
        c = t >> 32;
    }
    w[8] = (uint32_t)c;
    const uint32_t m = neg ? 0xffffffffu : 0u;
    uint64_t cc = neg ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint64_t t = (uint64_t)A.a[i] + (w[i] ^ m) + cc;
        A.a[i] = (uint32_t)t;
        cc = t >> 32;
    }
}
// the canonical residue of an accumulator within (-512 p, 512 p)
BN_INL fp wacc_reduce(const WAcc &A) {
    // D = A + 512 p in (0, 1024 p) < 2^264
    const uint32_t K512P[9] = {0xf9fa8e00u, 0x41182db0u, 0xe3951a78u, 0x02d522d0u, 0x02b0bb2fu, 0xa08b6d03u, 0x63405370u, 0xc89ce5c2u, 0x00000060u};
    uint32_t D[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint64_t t = (uint64_t)A.a[i] + K512P[i] + c;
        D[i] = (uint32_t)t;
        c = t >> 32;
    }
    // q = floor(D / p) or one less: floor(top64(D) * floor(2^64 / (top32(p) + 1)) / 2^64)
    const uint64_t top = (uint64_t)D[8] << 32 | D[7];
#if BN_PTX
    const uint32_t q = (uint32_t)__umul64hi(top, 22721021474ull);
#else
    const uint32_t q = (uint32_t)(((unsigned __int128)top * 22721021474ull) >> 64);
#endif
    // R = D - q p  in [0, 2p): fits eight words
    uint32_t R[8];
    uint64_t mc = 0;
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t t = (uint64_t)p_limb(i) * q + mc;
        mc = t >> 32;
        const int64_t dlt = (int64_t)D[i] - (int64_t)(uint32_t)t + bw;
        R[i] = (uint32_t)dlt;
        bw = dlt >> 32;      // 0 or -1
    }
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = R[i];
    fp_reduce_once(r);
    fp_reduce_once(r);
    return r;
}
BN_INL int wide_slot(const Env &E, uint32_t d) {
    int s = (int)(d & 1023u);
    if (d & 0x3000u) {
        if (d & 0x1000u) s += 3 * E.line;
        if (d & 0x2000u) s += 6 * E.var;
    }
    return s;
}
BN_INL fp wide_load(const WSlots M, const Env &E, uint32_t d, int half) {
    const int s = wide_slot(E, d);
    if (s >= W_CONST) return const_half(s - W_CONST, half != 0);
    return M.ldh(s, half);
}
// sum of the next n terms: +- [xi] coef * slot, this lane's half
BN_INL fp wide_gather(const WSlots M, const Env &E, WDesc &D, int n, bool odd) {
    WAcc A;
    wacc_zero(A);
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const uint32_t d = D.pop();
        const uint32_t k = (d >> 14) + 1u;
        const bool neg = (d & 0x400u) != 0;
        const fp v = wide_load(M, E, d, odd ? 1 : 0);
        if (d & 0x800u) {
            // xi * v: even lane 9 v0 - v1, odd lane 9 v1 + v0
            const fp o = wide_load(M, E, d, odd ? 0 : 1);
            wacc_mad(A, v, 9u * k, neg);
            wacc_mad(A, o, k, neg != !odd);
        } else {
            wacc_mad(A, v, k, neg);
        }
    }
    return wacc_reduce(A);
}

BN_NOINL void run(const WSlots M, const Env E, int routine) {
    const int lane = wide_lane(), unit = lane >> 1;
    const bool odd = (lane & 1) != 0;
    int b = M.dir(routine, 0);
    const int bend = b + M.dir(routine, 1);
#pragma unroll 1
    for (; b < bend; b++) {
        const int off = M.boff(b), cnt = M.boff(b + 1) - off;
        const bool active = unit < cnt;
        fp x, y, q, r;
        uint32_t w0 = 0;
        if (active) {
            uint32_t ins[8];
            M.ins(off + unit, ins);
            w0 = ins[0];
            WDesc D = {{ins[2], ins[3], ins[4], ins[5], ins[6], ins[7]}};
            const int op = (int)(w0 & 15u);
            const int nx = (int)((w0 >> 14) & 15u), ny = (int)((w0 >> 18) & 15u), nq = (int)((w0 >> 22) & 3u), nr = (int)((w0 >> 24) & 7u);
            x = wide_gather(M, E, D, nx, odd);
            if (op == W_MUL) y = ny ? wide_gather(M, E, D, ny, odd) : x;     // no Y terms: a squaring run as a general product
            else if (op == W_LIN && ny) y = wide_gather(M, E, D, ny, odd);    // linear step xi * (sum X) + (sum Y)
            else if (op == W_MULFP) y = wide_load(M, E, D.pop(), 0);     // the Fp scalar: c0 half, the same for both lanes
            if (nq) q = wide_gather(M, E, D, nq, odd);
            if (nr) r = wide_gather(M, E, D, nr, odd);
        }
        wide_sync();       // every operand of the bundle is in registers: results may now overwrite slots
        if (active) {
            const int op = (int)(w0 & 15u);
            int d = (int)((w0 >> 4) & 1023u);
            if ((w0 >> 31) & 1u) d += 6 * E.var;
            const int ny = (int)((w0 >> 18) & 15u), nq = (int)((w0 >> 22) & 3u), nr = (int)((w0 >> 24) & 7u), dbl = (int)((w0 >> 28) & 3u);
            const bool xi = (w0 >> 27) & 1u, conjx = (w0 >> 30) & 1u;
            fp p;
            if (op == W_LIN) {
                p = x;
                if (xi) p = wide_xi(p, odd);
                if (ny) fp_add(p, p, y);
#pragma unroll 1
                for (int k = 0; k < dbl; k++) fp_dbl(p, p);
                if (conjx) {
                    fp n;
                    fp_neg(n, p);
                    fp_select(p, odd, n, p);
                }
            } else if (op == W_INV) {
                // 1 / (x0 + x1 u) = (x0 - x1 u) / (x0^2 + x1^2)   (fp2.rs:289-303)
                fp px, n, ni;
                wide_xchg(px, x);
                fp_dot2_inl(n, x, x, px, px);
                fp_inv(ni, n);
                fp_mul_inl(p, x, ni);
                fp_neg(n, p);
                fp_select(p, odd, n, p);
            } else {
                if (conjx) {
                    fp n;
                    fp_neg(n, x);
                    fp_select(x, odd, n, x);
                }
                if (op == W_MUL) {
                    fp px, py, npy, u, w;
                    wide_xchg(px, x);
                    wide_xchg(py, y);
                    fp_neg(npy, py);
                    fp_select(u, odd, py, y);
                    fp_select(w, odd, y, npy);
                    fp_dot2_inl(p, x, u, px, w);
                } else {
                    if (op == W_SQR) {
                        fp px, s, dd;
                        wide_xchg(px, x);
                        fp_select(s, odd, px, x);
                        fp_sub(dd, x, px);
                        fp_select(y, odd, x, dd);
                        fp_add(x, s, px);
                    }
                    fp_mul_inl(p, x, y);
                }
                if (nq) fp_add(p, p, q);
                if (xi) p = wide_xi(p, odd);
                if (nr) fp_add(p, p, r);
#pragma unroll 1
                for (int k = 0; k < dbl; k++) fp_dbl(p, p);
            }
            M.sth(d, odd ? 1 : 0, p);
        }
        wide_sync();
    }
}
// the item's inputs into their slots (units 0..3: one scalar and one input each); absent pairs are passed as nullptr
BN_INL void wide_set_item(const WSlots M, const fp *p0, const fp2 *q0, const fp *p1, const fp2 *q1) {
    const int lane = wide_lane(), unit = lane >> 1;
    const bool odd = (lane & 1) != 0;
    if (unit < 4) {
        const fp *p = (unit & 2) ? p1 : p0;
        const fp2 *q = (unit & 2) ? q1 : q0;
        if (p) {            // the scalar as the Fp2 value (s, 0): MULFP reads the c0 half, a general product by the slot is the same
            fp v;
            fp_set_zero(v);
            if (!odd) v = p[unit & 1];
            M.sth(W_SC + unit, odd ? 1 : 0, v);
        }
        if (q) M.sth(W_IN + unit, odd ? 1 : 0, global_half(q + (unit & 1), odd));
    }
    wide_sync();
}
BN_INL void load_f12(const WSlots M, const fp12 &F, int first_slot = 0) {
    const int lane = wide_lane(), unit = lane >> 1;
    const fp2 *c = reinterpret_cast<const fp2 *>(&F);
    if (unit < 6) M.sth(first_slot + unit, lane & 1, global_half(&c[unit], (lane & 1) != 0));
    wide_sync();
}
BN_INL void store_f12(fp12 &F, const WSlots M) {
    const int lane = wide_lane(), unit = lane >> 1;
    fp2 *c = reinterpret_cast<fp2 *>(&F);
    if (unit < 6) {
        fp *p = (lane & 1) ? &c[unit].c1 : &c[unit].c0;
        *p = M.ldh(unit, lane & 1);
    }
}

// ---- drivers: the loops over NAF / window digits ---------------------------------------------------------------------
enum Family : int { FAM_M1 = 0, FAM_M2F = 1, FAM_M2V = 2 };

// Miller loop of one item (pairings.rs:808-857, 862-886 with the BN254 loop of SURVEY.md 0.5); result in slots 0..5
template <class S>
BN_INL void miller(const S M, Env &E, int fam) {
    const int r_init = fam == FAM_M2V ? R_M2V_INIT : R_M2F_INIT;
    const int r_dbl = fam == FAM_M1 ? R_M1_DBL : (fam == FAM_M2F ? R_M2F_DBL : R_M2V_DBL);
    const int r_addp = fam == FAM_M1 ? R_M1_ADDP : (fam == FAM_M2F ? R_M2F_ADDP : R_M2V_ADDP);
    const int r_addn = fam == FAM_M1 ? R_M1_ADDN : (fam == FAM_M2F ? R_M2F_ADDN : R_M2V_ADDN);
    const int r_frob = fam == FAM_M1 ? R_M1_FROB : (fam == FAM_M2F ? R_M2F_FROB : R_M2V_FROB);
    E.line = 0;
    run(M, E, r_init);
#pragma unroll 1
    for (int s = 0; s < BN254_ATE_NAF_LEN; s++) {
        const int d = k::ATE_NAF[s];
        run(M, E, r_dbl);
        E.line++;
        if (d) {
            run(M, E, d > 0 ? r_addp : r_addn);
            E.line++;
        }
    }
    run(M, E, r_frob);
}

// scratch variables of the final exponentiation
enum : int { V_R = 0, V_T = 1, V_A = 2, V_B = 3, V_C = 4, V_PW = 5 };
template <class S>
BN_INL void fe(const S M, Env &E, int routine) { run(M, E, routine); }
template <class S>
BN_INL void fe_var(const S M, Env &E, int routine, int var) {
    E.var = var;
    run(M, E, routine);
}
// F <- conj(F^x): width-4 signed window over f, f^3, f^5, f^7 kept in scratch (pairing.cuh: cyclotomic_exp_neg_x)
template <class S>
BN_NOINL void fe_exp_neg_x(const S M, Env &E) {
    fe_var(M, E, R_FE_STF, V_PW);
    fe(M, E, R_FE_MOVGF);
    fe(M, E, R_FE_CYCSQR);
    fe_var(M, E, R_FE_STF, V_T);
#pragma unroll 1
    for (int i = 1; i < 4; i++) {
        fe_var(M, E, R_FE_LDF, V_PW + i - 1);
        fe_var(M, E, R_FE_LDG, V_T);
        fe(M, E, R_FE_MUL);
        fe_var(M, E, R_FE_STF, V_PW + i);
    }
    fe_var(M, E, R_FE_LDF, V_PW + ((k::X_WNAF4[0] - 1) >> 1));
#pragma unroll 1
    for (int i = 1; i < BN254_X_WNAF_LEN; i++) {
        fe(M, E, R_FE_CYCSQR);
        const int d = k::X_WNAF4[i];
        if (d > 0) {
            fe_var(M, E, R_FE_LDG, V_PW + ((d - 1) >> 1));
            fe(M, E, R_FE_MUL);
        } else if (d < 0) {
            fe(M, E, R_FE_CONJ);
            fe_var(M, E, R_FE_LDG, V_PW + ((-d - 1) >> 1));
            fe(M, E, R_FE_MUL);
            fe(M, E, R_FE_CONJ);
        }
    }
    fe(M, E, R_FE_CONJ);
}
// pairings.rs:50-178 role (easy part, then Fuentes-Castaneda): input and result in slots 0..5.  The step list is the one
// tools/gen_vm_programs.py::sim_final_exp checks against the big-integer model.
template <class S>
BN_INL void final_exp(const S M, Env &E) {
    fe_var(M, E, R_FE_STF, V_A);
    fe(M, E, R_FE_INV);
    fe_var(M, E, R_FE_STF, V_T);
    fe_var(M, E, R_FE_LDF, V_A);
    fe(M, E, R_FE_CONJ);
    fe_var(M, E, R_FE_LDG, V_T);
    fe(M, E, R_FE_MUL);                  // conj(f) / f
    fe_var(M, E, R_FE_STF, V_T);
    fe(M, E, R_FE_FROB2);
    fe_var(M, E, R_FE_LDG, V_T);
    fe(M, E, R_FE_MUL);
    fe_var(M, E, R_FE_STF, V_R);         // r = f^((p^6 - 1)(p^2 + 1))
    fe_exp_neg_x(M, E);                  // y0
    fe(M, E, R_FE_CYCSQR);               // y1
    fe_var(M, E, R_FE_STF, V_A);
    fe(M, E, R_FE_CYCSQR);               // y2
    fe_var(M, E, R_FE_LDG, V_A);
    fe(M, E, R_FE_MUL);                  // y3
    fe_var(M, E, R_FE_STF, V_B);
    fe_exp_neg_x(M, E);                  // y4
    fe_var(M, E, R_FE_STF, V_C);
    fe(M, E, R_FE_CYCSQR);               // y5
    fe_exp_neg_x(M, E);                  // y6
    fe(M, E, R_FE_CONJ);
    fe_var(M, E, R_FE_LDG, V_C);
    fe(M, E, R_FE_MUL);                  // y7
    fe_var(M, E, R_FE_STF, V_T);
    fe_var(M, E, R_FE_LDF, V_B);
    fe(M, E, R_FE_CONJ);
    fe_var(M, E, R_FE_LDG, V_T);
    fe(M, E, R_FE_MUL);                  // y8
    fe_var(M, E, R_FE_STF, V_B);
    fe_var(M, E, R_FE_LDG, V_A);
    fe(M, E, R_FE_MUL);                  // y9
    fe_var(M, E, R_FE_STF, V_A);
    fe_var(M, E, R_FE_LDF, V_B);
    fe_var(M, E, R_FE_LDG, V_C);
    fe(M, E, R_FE_MUL);                  // y10
    fe_var(M, E, R_FE_LDG, V_R);
    fe(M, E, R_FE_MUL);                  // y11
    fe_var(M, E, R_FE_STF, V_T);
    fe_var(M, E, R_FE_LDF, V_A);
    fe(M, E, R_FE_FROB1);                // y12
    fe_var(M, E, R_FE_LDG, V_T);
    fe(M, E, R_FE_MUL);                  // y13
    fe_var(M, E, R_FE_STF, V_T);
    fe_var(M, E, R_FE_LDF, V_B);
    fe(M, E, R_FE_FROB2);
    fe_var(M, E, R_FE_LDG, V_T);
    fe(M, E, R_FE_MUL);                  // y14
    fe_var(M, E, R_FE_STF, V_T);
    fe_var(M, E, R_FE_LDF, V_R);
    fe(M, E, R_FE_CONJ);
    fe_var(M, E, R_FE_LDG, V_A);
    fe(M, E, R_FE_MUL);                  // y15
    fe(M, E, R_FE_FROB3);
    fe_var(M, E, R_FE_LDG, V_T);
    fe(M, E, R_FE_MUL);
}

// slots 0..5 <-> the item's one-thread Fp12 record in HBM (each lane moves its half)
BN_INL void load_f12(const Slots M, const fp12 &F, bool odd) {
    const fp2 *c = reinterpret_cast<const fp2 *>(&F);
#pragma unroll 1
    for (int i = 0; i < 6; i++) M.st(i, global_half(&c[i], odd));
}
BN_INL void store_f12(fp12 &F, const Slots M, bool odd) {
    fp2 *c = reinterpret_cast<fp2 *>(&F);
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        fp *p = odd ? &c[i].c1 : &c[i].c0;
        *p = M.ld(i);
    }
}

}  // namespace vm
}  // namespace bn
