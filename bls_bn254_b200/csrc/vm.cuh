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

// ---- drivers: the loops over NAF / window digits ---------------------------------------------------------------------
enum Family : int { FAM_M1 = 0, FAM_M2F = 1, FAM_M2V = 2 };

// Miller loop of one item (pairings.rs:808-857, 862-886 with the BN254 loop of SURVEY.md 0.5); result in slots 0..5
BN_INL void miller(const Slots M, Env &E, int fam) {
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
BN_INL void fe(const Slots M, Env &E, int routine) { run(M, E, routine); }
BN_INL void fe_var(const Slots M, Env &E, int routine, int var) {
    E.var = var;
    run(M, E, routine);
}
// F <- conj(F^x): width-4 signed window over f, f^3, f^5, f^7 kept in scratch (pairing.cuh: cyclotomic_exp_neg_x)
BN_NOINL void fe_exp_neg_x(const Slots M, Env &E) {
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
BN_INL void final_exp(const Slots M, Env &E) {
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
