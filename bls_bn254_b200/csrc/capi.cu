// capi.cu -- the C ABI of libbls_bn254_b200.so (include/bls_bn254_b200.h): context, scratch
// memory, chunked host<->device staging, sharding over the ctx's devices, kernel launches.
// There is no CPU path: without a CUDA device bn254_ctx_create fails with BN254_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bls_bn254_b200.h"
#include "kernels.cuh"
#include "nccl_dyn.h"

using namespace bn;

namespace {

enum Scratch {
    S_IN_A, S_IN_B, S_IN_C, S_MSGS, S_OFF, S_DST, S_OUT, S_STATUS,
    S_P0, S_FLP0, S_P1, S_FLP1, S_Q0, S_FLQ0, S_Q1, S_FLQ1, S_F, S_FLF, S_F2, S_BAD, S_TMP, S_FLTMP, S_CONST,
    S_RB_S0, S_RB_S1, S_RB_S2, S_RB_FA, S_RB_FB, S_RB_FS, S_RB_GP, S_RB_GQ, S_RB_FLP, S_RB_FLQ, S_RB_BAD, S_RB_SEED,
    // second set of staging buffers: the chunk pipeline copies chunk k+1 in while the kernels of chunk k read the first set
    S_IN_A2, S_IN_B2, S_IN_C2, S_MSGS2, S_OFF2, S_OUT2, S_STATUS2,
    S_AGG_SEND, S_AGG_RECV,   // agg_rec: this rank's record / the gathered records of all ranks
    S_VM_SCRATCH,             // per-item scratch variables of the slot-machine final exponentiation
    S_RB_T, S_RB_T2, S_RB_BAD2,   // per staging set: group products / bad flags handed to the overlapped tail of the randomised path
    S_COUNT
};
// staging set s (0 / 1) of the chunk pipeline
inline Scratch stage_buf(Scratch first, int s) {
    if (!s) return first;
    switch (first) {
        case S_IN_A: return S_IN_A2;
        case S_IN_B: return S_IN_B2;
        case S_IN_C: return S_IN_C2;
        case S_MSGS: return S_MSGS2;
        case S_OFF: return S_OFF2;
        case S_OUT: return S_OUT2;
        case S_STATUS: return S_STATUS2;
        case S_RB_T: return S_RB_T2;
        case S_RB_BAD: return S_RB_BAD2;
        default: return first;
    }
}

struct Buf { void *p = nullptr; size_t cap = 0; };

struct ProfRec { const char *name; size_t items; cudaEvent_t e0, e1; };

struct Dev {
    int id = 0;
    cudaStream_t st = nullptr;
    Buf bufs[S_COUNT];
    uint64_t launches = 0;
    std::string err;
    int rc = 0;
    bool prof = false;            // record a CUDA event pair around every kernel launch
    std::vector<ProfRec> recs;
    // A verification batch can be cut into `nsub` slices whose pipelines (decode, hash, Miller, final
    // exponentiation, compare) run on separate streams, so that one slice's kernels take the CTA
    // slots another leaves idle (a 2^16 batch fills 86 % of one wave).  Measured on B200 it LOSES:
    // 36.2 ms unsliced vs 40.0-40.7 ms with 2-8 slices (different kernels sharing an SM thrash the
    // instruction cache and L1), so the default is 1; BN254_SPLIT=k enables it.
    static const int MAXSUB = 8;
    cudaStream_t sub[MAXSUB] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[MAXSUB] = {};
    cudaStream_t cur = nullptr;   // the stream LAUNCH uses (d.st outside a split region)
    // High-priority side stream: the latency-bound per-group tail of the randomised batch verification (a few
    // hundred threads running whole Miller loops) is queued here so that its CTAs are placed as soon as a slot
    // frees up while the full-size kernels of the main stream keep the machine busy.
    cudaStream_t hi = nullptr;
    cudaEvent_t ev_hi_fork = nullptr, ev_hi_join = nullptr;
    // Second high-priority stream: the per-group TAIL of a randomised chunk (one final exponentiation per group on a few
    // thousand threads, ~4 ms of latency that fills 2 % of the machine) runs here, next to the main kernels of the NEXT chunk.
    cudaStream_t tail = nullptr;
    cudaEvent_t ev_tail[2] = {};
    int nsub = 1;
    size_t lane_wave_threads = 0;   // threads of one full wave of the lane-pair kernels on this device
    size_t sms = 1;
    size_t sk_msm_slices = 0;       // tuning: slices per segment of the bucket kernels (0 = chosen per chunk)
    size_t sk_msm_min = 128;        // same-key path: average tuples per key from which the sums run as bucket-method MSMs
    // Shape of the pairing kernels: 0 = picked per launch (lane pairs while 2n threads fit one wave), 1 = one thread per
    // pairing, 2 = two lanes per pairing.  BN254_LANES / bn254_ctx_set_option("lanes") force a shape so that the parity
    // suite meets BOTH shapes at every size (the big-batch shape is what a 2^24-tuple job runs).
    int shape = 0;
    // 3 = the slot machine, 4 = the WIDE slot machine (one warp per item: the latency shape).  In shape 0 a launch of at most
    // wide_max items uses the wide kernels ("wide_max" option / BN254_WIDE_MAX; 0 = never).  Default: about ten warps per SM,
    // below which a warp per item beats a lane pair per item (measured per verification batch: 256 items 4.1 against 8.7 ms,
    // 1024 items 4.9 against 8.8 ms, 2048 items 9.0 against 8.9 ms, 3072 items 11.8 against 9.0 ms).
    size_t wide_max = 1536;
    // Chunk pipeline of the host-buffer entry points: copies run on `cp`, kernels on `st`; two staging sets.
    cudaStream_t cp = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_in2[2] = {}, ev_out[2] = {};   // ev_in2: the keys / signatures of a chunk have arrived (hashing only waits for ev_in: the messages)
    void *pin[2] = {};              // pinned result staging (used when the caller's result buffer is pageable)
    size_t pin_cap[2] = {};
    // NCCL communicator of the aggregate path (one rank per device; null on a lone device without bn254_comm_init_rank)
    bn_nccl::comm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

}  // namespace

struct bn254_ctx {
    std::vector<Dev> devs;
    std::string err;
    size_t chunk = (size_t)1 << 18;  // items staged per host<->device round
    bool chunk_is_default = true;    // not set through BN254_CHUNK / bn254_ctx_set_option
    // Exact per-tuple verification streams chunks of exactly one wave of the lane-pair kernels (66 304 tuples on 148 SMs = one
    // 896-thread CTA on every SM: 2.35 M verifications/s per launch against 2.32 M for 2^18-tuple launches); with the
    // double-buffered pipeline the chunk size does not matter for the copies.  The randomised path keeps 2^18.
    size_t verify_chunk(const Dev &d) const { return chunk_is_default && d.shape != 1 ? std::min(chunk, d.lane_wave_threads / 2) : chunk; }
};

// G2Prepared for a set of public keys: per device of the ctx the 88 line triples of every key (16 896 bytes per key) and one
// flag byte per key (undecodable / identity / failed the requested checks -> every tuple that names the key is rejected)
struct bn254_g2_prepared {
    bn254_ctx *ctx = nullptr;
    size_t n = 0;
    std::vector<void *> tab, fl, pts;    // per device: line tables, key flags, the decoded keys themselves (g2m)
};

namespace {

#define CK(d, call)                                                                                                 \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) {                                                                                    \
            (d).err = std::string(#call) + ": " + cudaGetErrorString(e_);                                          \
            (d).rc = BN254_ERR_CUDA_BASE - (int)e_;                                                                 \
            return (d).rc;                                                                                          \
        }                                                                                                           \
    } while (0)

int ensure(Dev &d, Scratch s, size_t bytes, void **out) {
    Buf &b = d.bufs[s];
    if (bytes == 0) bytes = 16;
    if (b.cap < bytes) {
        if (b.p) {
            CK(d, cudaStreamSynchronize(d.st));
            if (d.cp) CK(d, cudaStreamSynchronize(d.cp));
            CK(d, cudaFree(b.p));
            b.p = nullptr;
            b.cap = 0;
        }
        size_t cap = bytes + bytes / 8 + 256;
        CK(d, cudaMalloc(&b.p, cap));
        b.cap = cap;
    }
    *out = b.p;
    return 0;
}
template <class T>
int ensure_t(Dev &d, Scratch s, size_t count, T **out) {
    void *p = nullptr;
    int rc = ensure(d, s, count * sizeof(T), &p);
    *out = (T *)p;
    return rc;
}

inline unsigned blocks(size_t n, unsigned tpb) { return (unsigned)((n + tpb - 1) / tpb); }

#define BN_STR2(x) #x
#define BN_STR(x) BN_STR2(x)
// per-launch timing: the event pair is owned by d.recs from the moment it exists, so no error path leaks it
int prof_begin(Dev &d, const char *name, size_t n) {
    if (!d.prof) return 0;
    ProfRec pr = {name, n, nullptr, nullptr};
    d.recs.push_back(pr);
    ProfRec &r = d.recs.back();
    CK(d, cudaEventCreate(&r.e0));
    CK(d, cudaEventCreate(&r.e1));
    CK(d, cudaEventRecord(r.e0, d.cur));
    return 0;
}
int prof_end(Dev &d) {
    if (!d.prof) return 0;
    CK(d, cudaEventRecord(d.recs.back().e1, d.cur));
    return 0;
}
// the launch itself failed, or an earlier asynchronous fault surfaced here: say which launch reported it
int launch_check(Dev &d, const char *kernel) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    d.err = std::string("at the launch of ") + kernel + " (asynchronous errors of earlier kernels surface here too): " + cudaGetErrorString(e);
    d.rc = BN254_ERR_CUDA_BASE - (int)e;
    return d.rc;
}
#define LAUNCH(d, kern, n, tpb, ...)                                             \
    do {                                                                         \
        if ((n) > 0) {                                                           \
            if (int rcp_ = prof_begin(d, BN_STR(kern), (size_t)(n))) return rcp_; \
            kern<<<blocks((n), (tpb)), (tpb), 0, (d).cur>>>(__VA_ARGS__);         \
            (d).launches++;                                                      \
            if (int rcp_ = prof_end(d)) return rcp_;                             \
            if (int rcl_ = launch_check(d, BN_STR(kern))) return rcl_;           \
        }                                                                        \
    } while (0)

// Pairing kernels come in bit-identical shapes: one thread per pairing, and two lanes per pairing (lane-split Fp2, tower.cuh:
// half the per-thread state, twice the warps).  d.shape = 0 (default) runs the lane pair at every size: it wins below one
// wave because one thread per item leaves the machine under-occupied (2^14: +19 %, 2^16: +8 %) and, measured on the final
// build, also beyond (2^18 items per launch: two-pair loop +5 %, one-pair loop +10 %, final exponentiation +5 %, twin loop
// +4 %; 2^20: verification +5.5 %).  1 / 2 force a shape (BN254_LANES, bn254_ctx_set_option) so that the parity suite meets
// the one-thread kernels too.
// CTA size of a lane-pair launch of `threads` threads: one CTA per SM (a multiple of 32 threads, at most BN_TPB_L2) once the
// launch has at least 64 threads per SM, else 64-thread CTAs on as many SMs as there are CTAs.
inline unsigned lane_tpb(const Dev &d, size_t threads) {
    const size_t per_sm = (threads + d.sms - 1) / d.sms;
    if (per_sm >= BN_TPB_L2) return BN_TPB_L2;
    if (per_sm <= 64) return 64;
    return (unsigned)((per_sm + 31) / 32 * 32);
}
#define LAUNCH_PAIR(d, kscalar, klane, n, ...)                                              \
    do {                                                                                    \
        if ((d).shape != 1)                                                                 \
            LAUNCH(d, klane, 2 * (size_t)(n), lane_tpb(d, 2 * (size_t)(n)), __VA_ARGS__);   \
        else                                                                                \
            LAUNCH(d, kscalar, (size_t)(n), BN_TPB_HEAVY, __VA_ARGS__);                     \
    } while (0)

// The slot-machine shape (d.shape == 3, vm.cuh): a lane pair per item like the lane kernels, one warp per CTA, dynamic shared
// memory = slots x 32 bytes x BN_VM_TPB.
#define LAUNCH_VM(d, kern, n, nslot, ...)                                                                   \
    do {                                                                                                    \
        if ((n) > 0) {                                                                                      \
            if (int rcp_ = prof_begin(d, BN_STR(kern), 2 * (size_t)(n))) return rcp_;                       \
            kern<<<blocks(2 * (size_t)(n), BN_VM_TPB), BN_VM_TPB, BN_VM_SMEM(nslot), (d).cur>>>(__VA_ARGS__); \
            (d).launches++;                                                                                 \
            if (int rcp_ = prof_end(d)) return rcp_;                                                        \
            if (int rcl_ = launch_check(d, BN_STR(kern))) return rcl_;                                      \
        }                                                                                                   \
    } while (0)

// The wide slot machine (d.shape == 4, or automatically for n <= d.wide_max): one WARP per item (vm.cuh: run on WSlots).
#define LAUNCH_W(d, kern, r_lo, r_hi, n, ...)                                                      \
    do {                                                                                           \
        if ((n) > 0) {                                                                             \
            const size_t smem_ = wide_smem(vm::r_lo, vm::r_hi);                                    \
            static std::atomic<unsigned> attr_set_{0};   /* per kernel and device: opt in to > 48 KB of shared memory */ \
            if (!((attr_set_.load() >> (d).id) & 1u)) {                                            \
                CK(d, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_)); \
                attr_set_.fetch_or(1u << (d).id);                                                  \
            }                                                                                      \
            if (int rcp_ = prof_begin(d, BN_STR(kern), (size_t)(n))) return rcp_;                  \
            kern<<<blocks((n), BN_W_WARPS), 32 * BN_W_WARPS, smem_, (d).cur>>>(__VA_ARGS__);       \
            (d).launches++;                                                                        \
            if (int rcp_ = prof_end(d)) return rcp_;                                               \
            if (int rcl_ = launch_check(d, BN_STR(kern))) return rcl_;                             \
        }                                                                                          \
    } while (0)
// shared memory of a wide-machine CTA running routines [r_lo, r_hi] (host copy of the generated tables)
inline size_t wide_smem(int r_lo, int r_hi) {
    static uint16_t dir[vm::R_COUNT][2];
    static uint16_t boff[sizeof(vm::W_BOFF) / 2];
    static std::once_flag once;      // a multi-device ctx launches from one host thread per device
    std::call_once(once, [] {
        cudaMemcpyFromSymbol(dir, vm::W_DIR, sizeof dir);
        cudaMemcpyFromSymbol(boff, vm::W_BOFF, sizeof boff);
    });
    const int i_lo = boff[dir[r_lo][0]], i_hi = boff[dir[r_hi][0] + dir[r_hi][1]];
    return vm::wide_smem_bytes(r_lo, r_hi, BN_W_WARPS, i_lo, i_hi);
}
inline bool use_wide(const Dev &d, size_t n) { return d.shape == 4 || (d.shape == 0 && n <= d.wide_max); }

// ---- the pairing kernels behind one call each: the shape (one thread / lane pair / slot machine) is picked here -----------
int launch_miller2_fixed(Dev &d, size_t n, fp12 *F, uint8_t *flf, const g1m *P0, const uint8_t *flp0, const g1m *P1, const uint8_t *flp1,
                         const g2m *Q1, const uint8_t *flq1) {
    if (use_wide(d, n)) {
        LAUNCH_W(d, k_w_miller2_fixed, R_M2F_INIT, R_M2F_FROB, n, F, flf, P0, flp0, P1, flp1, Q1, flq1, n);
        LAUNCH(d, k_vm_fixup_miller2_fixed, n, BN_TPB_HEAVY, F, P0, flp0, P1, flp1, Q1, flq1, n);
        return 0;
    }
    if (d.shape == 3) {
        LAUNCH_VM(d, k_vm_miller2_fixed, n, vm::NSLOT_M1, F, flf, P0, flp0, P1, flp1, Q1, flq1, n);
        LAUNCH(d, k_vm_fixup_miller2_fixed, n, BN_TPB_HEAVY, F, P0, flp0, P1, flp1, Q1, flq1, n);
        return 0;
    }
    LAUNCH_PAIR(d, k_miller2_fixed, k_miller2_fixed_l2, n, F, flf, P0, flp0, P1, flp1, Q1, flq1, n);
    return 0;
}
int launch_miller2(Dev &d, size_t n, fp12 *F, uint8_t *flf, const g1m *P0, const uint8_t *flp0, size_t p0_stride, const g2m *Q0, const uint8_t *flq0,
                   const g1m *P1, const uint8_t *flp1, const g2m *Q1, const uint8_t *flq1) {
    if (use_wide(d, n)) {
        LAUNCH_W(d, k_w_miller2, R_M2V_INIT, R_M2V_FROB, n, F, flf, P0, flp0, p0_stride, Q0, flq0, P1, flp1, Q1, flq1, n, (size_t)1);
        LAUNCH(d, k_vm_fixup_miller2, n, BN_TPB_HEAVY, F, P0, flp0, p0_stride, Q0, flq0, P1, flp1, Q1, flq1, n, (size_t)1, n);
        return 0;
    }
    if (d.shape == 3) {
        LAUNCH_VM(d, k_vm_miller2, n, vm::NSLOT_M2V, F, flf, P0, flp0, p0_stride, Q0, flq0, P1, flp1, Q1, flq1, n, (size_t)1);
        LAUNCH(d, k_vm_fixup_miller2, n, BN_TPB_HEAVY, F, P0, flp0, p0_stride, Q0, flq0, P1, flp1, Q1, flq1, n, (size_t)1, n);
        return 0;
    }
    LAUNCH_PAIR(d, k_miller2, k_miller2_l2, n, F, flf, P0, flp0, p0_stride, Q0, flq0, P1, flp1, Q1, flq1, n);
    return 0;
}
int launch_miller1(Dev &d, size_t n, fp12 *F, uint8_t *flf, const g1m *P, const uint8_t *flp, const g2m *Q, const uint8_t *flq, int *bad) {
    if (use_wide(d, n)) {
        LAUNCH_W(d, k_w_miller1, R_M2F_INIT, R_M1_FROB, n, F, flf, P, flp, Q, flq, n, bad);
        return 0;
    }
    if (d.shape == 3) {
        LAUNCH_VM(d, k_vm_miller1, n, vm::NSLOT_M1, F, flf, P, flp, Q, flq, n, bad);
        return 0;
    }
    LAUNCH_PAIR(d, k_miller1, k_miller1_l2, n, F, flf, P, flp, Q, flq, n, bad);
    return 0;
}
// F[t] = ML(P[2t], Q[2t]) * ML(P[2t+1], Q[2t+1]) for t < ceil(ntuples / 2)
int launch_miller_twin(Dev &d, size_t ntuples, fp12 *F, const g1m *P, const uint8_t *flp, const g2m *Q, const uint8_t *flq, int *bad) {
    const size_t half = (ntuples + 1) / 2;
    if (use_wide(d, half)) {
        LAUNCH_W(d, k_w_miller2, R_M2V_INIT, R_M2V_FROB, half, F, (uint8_t *)nullptr, P, flp, (size_t)1, Q, flq, P + 1, flp + 1, Q + 1, flq + 1, half, (size_t)2);
        LAUNCH(d, k_vm_fixup_miller2, half, BN_TPB_HEAVY, F, P, flp, (size_t)1, Q, flq, P + 1, flp + 1, Q + 1, flq + 1, half, (size_t)2, ntuples);
        if (bad) LAUNCH(d, k_or_flags, ntuples, BN_TPB_LIGHT, bad, flp, ntuples);
        if (bad) LAUNCH(d, k_or_flags, ntuples, BN_TPB_LIGHT, bad, flq, ntuples);
        return 0;
    }
    if (d.shape == 3) {
        LAUNCH_VM(d, k_vm_miller2, half, vm::NSLOT_M2V, F, (uint8_t *)nullptr, P, flp, (size_t)1, Q, flq, P + 1, flp + 1, Q + 1, flq + 1, half, (size_t)2);
        LAUNCH(d, k_vm_fixup_miller2, half, BN_TPB_HEAVY, F, P, flp, (size_t)1, Q, flq, P + 1, flp + 1, Q + 1, flq + 1, half, (size_t)2, ntuples);
        if (bad) LAUNCH(d, k_or_flags, ntuples, BN_TPB_LIGHT, bad, flp, ntuples);
        if (bad) LAUNCH(d, k_or_flags, ntuples, BN_TPB_LIGHT, bad, flq, ntuples);
        return 0;
    }
    LAUNCH_PAIR(d, k_miller_twin, k_miller_twin_l2, half, F, P, flp, Q, flq, ntuples, bad);
    return 0;
}
int launch_final_exp(Dev &d, fp12 *F, size_t n) {
    if (use_wide(d, n)) {
        LAUNCH_W(d, k_w_final_exp, R_FE_CYCSQR, R_FE_STF, n, F, n);
        return 0;
    }
    if (d.shape == 3) {
        fp2 *scratch;
        if (int rc = ensure_t(d, S_VM_SCRATCH, (size_t)vm::FEXP_SCRATCH_FP2 * n, &scratch)) return rc;
        LAUNCH_VM(d, k_vm_final_exp, n, vm::NSLOT_FEXP, F, n, scratch);
        return 0;
    }
    LAUNCH_PAIR(d, k_final_exp, k_final_exp_l2, n, F, n);
    return 0;
}

// ---- device-level building blocks (all pointers are device pointers on d) ------------------------

// DST: copies / pre-hashes into scratch when needed.  `dst_on_host` tells where dst lives.
int stage_dst(Dev &d, const uint8_t *dst, size_t dst_len, bool dst_on_host, const uint8_t **d_dst, uint32_t *d_len) {
    uint8_t *buf;
    if (int rc = ensure_t(d, S_DST, dst_len + 64, &buf)) return rc;
    const uint8_t *src = dst;
    if (dst_on_host) {
        if (dst_len) CK(d, cudaMemcpyAsync(buf + 32, dst, dst_len, cudaMemcpyHostToDevice, d.st));
        src = buf + 32;
    }
    if (dst_len > 255) {  // RFC 9380 5.3.3
        LAUNCH(d, k_oversize_dst, 1, 32, buf, src, dst_len);
        *d_dst = buf;
        *d_len = 32;
    } else {
        *d_dst = src;
        *d_len = (uint32_t)dst_len;
    }
    return 0;
}

int dev_hash_g1(Dev &d, g1m *out, uint8_t *fl, const uint8_t *msgs, const uint64_t *off, uint64_t base, size_t n, const uint8_t *dst,
                uint32_t dst_len, int count) {
    LAUNCH(d, k_hash_g1, n, BN_TPB_HASH1, out, fl, msgs, off, base, n, dst, dst_len, count);
    return 0;
}
int dev_hash_g2(Dev &d, g2m *out, uint8_t *fl, const uint8_t *msgs, const uint64_t *off, uint64_t base, size_t n, const uint8_t *dst,
                uint32_t dst_len, int count) {
    LAUNCH(d, k_hash_g2, n, BN_TPB_HASH2, out, fl, msgs, off, base, n, dst, dst_len, count);
    return 0;
}

// hash / encode to wire bytes
int dev_hash_wire(Dev &d, int group, const uint8_t *msgs, const uint64_t *off, uint64_t base, size_t n, const uint8_t *dst,
                  uint32_t dst_len, int count, uint8_t *out, uint8_t *status) {
    uint8_t *fl;
    if (int rc = ensure_t(d, S_FLP1, n, &fl)) return rc;
    if (group == 1) {
        g1m *P;
        if (int rc = ensure_t(d, S_P1, n, &P)) return rc;
        if (int rc = dev_hash_g1(d, P, fl, msgs, off, base, n, dst, dst_len, count)) return rc;
        LAUNCH(d, k_g1_encode, n, BN_TPB_LIGHT, out, status, P, fl, n);
    } else {
        g2m *Q;
        if (int rc = ensure_t(d, S_Q1, n, &Q)) return rc;
        if (int rc = dev_hash_g2(d, Q, fl, msgs, off, base, n, dst, dst_len, count)) return rc;
        LAUNCH(d, k_g2_encode, n, BN_TPB_LIGHT, out, status, Q, fl, n);
    }
    return 0;
}

// mode 0: pairing, 1: Miller loop only
int dev_pairing(Dev &d, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *out, uint8_t *status, int mode) {
    g1m *P; g2m *Q; uint8_t *flp, *flq, *flf; fp12 *F;
    if (int rc = ensure_t(d, S_P0, n, &P)) return rc;
    if (int rc = ensure_t(d, S_Q0, n, &Q)) return rc;
    if (int rc = ensure_t(d, S_FLP0, n, &flp)) return rc;
    if (int rc = ensure_t(d, S_FLQ0, n, &flq)) return rc;
    if (int rc = ensure_t(d, S_FLF, n, &flf)) return rc;
    if (int rc = ensure_t(d, S_F, n, &F)) return rc;
    LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P, flp, g1, n, 0);
    LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, Q, flq, g2, n, 0);
    if (int rc = launch_miller1(d, n, F, flf, P, flp, Q, flq, (int *)nullptr)) return rc;
    if (mode == 0)
        if (int rc = launch_final_exp(d, F, n)) return rc;
    LAUNCH(d, k_fp12_encode, n, BN_TPB_LIGHT, out, status, F, flf, n);
    return 0;
}

int dev_final_exp(Dev &d, const uint8_t *ml, size_t n, uint8_t *out, uint8_t *status) {
    fp12 *F; uint8_t *flf;
    if (int rc = ensure_t(d, S_F, n, &F)) return rc;
    if (int rc = ensure_t(d, S_FLF, n, &flf)) return rc;
    LAUNCH(d, k_fp12_decode, n, BN_TPB_LIGHT, F, flf, ml, n);
    if (int rc = launch_final_exp(d, F, n)) return rc;
    LAUNCH(d, k_fp12_encode, n, BN_TPB_LIGHT, out, status, F, flf, n);
    return 0;
}

// product of n Fp12 records (internal form) -> *result points at one record: `final_dst` when given (the last pass of
// the tree then writes straight into it, e.g. the NCCL send record), else a scratch record
// Fan-in per pass: 16 while a pass still fills the machine, 4 below that -- a pass is then a latency-bound chain of fan - 1
// dependent Fp12 products per thread (~23 us each), so 2^16 partials cost 8 x 3 products instead of 4 x 15.
int dev_fp12_product(Dev &d, fp12 *F, size_t n, fp12 **result, fp12 *final_dst = nullptr) {
    auto fan_of = [](size_t cur) -> size_t { return cur > ((size_t)1 << 18) ? 16 : 4; };
    fp12 *A, *B;
    size_t m0 = std::max<size_t>(1, (n + fan_of(n) - 1) / fan_of(n));
    if (int rc = ensure_t(d, S_F2, m0 + (m0 + 3) / 4 + 2, &A)) return rc;
    B = A + m0;
    const fp12 *src = F;
    size_t cur = n;
    fp12 *dstp = A;
    while (true) {
        const size_t fan = fan_of(cur);
        size_t m = std::max<size_t>(1, (cur + fan - 1) / fan);
        if (m == 1 && final_dst) dstp = final_dst;
        LAUNCH(d, k_fp12_reduce, m, BN_TPB_HEAVY, dstp, src, cur, m);
        if (m == 1) break;
        src = dstp;
        cur = m;
        dstp = (dstp == A) ? B : A;
    }
    *result = dstp;
    return 0;
}

// -g1 = (1, -2) in Montgomery form + a zero flag byte, filled once per device by a kernel on the launch stream
int neg_g1_const(Dev &d, const g1m **pt, const uint8_t **fl) {
    Buf &b = d.bufs[S_CONST];
    if (!b.p) {
        void *p = nullptr;
        if (int rc = ensure(d, S_CONST, sizeof(g1m) + 16, &p)) return rc;
        LAUNCH(d, k_fill_neg_g1, 1, BN_TPB_LIGHT, (g1m *)p, (uint8_t *)p + sizeof(g1m), (size_t)1);
    }
    *fl = (const uint8_t *)b.p + sizeof(g1m);
    *pt = (const g1m *)b.p;
    return 0;
}

// verification of items [lo, hi) of one chunk on the current stream (all arrays are chunk-sized)
struct VerifyBufs {
    fp12 *F; uint8_t *flf; g1m *P0, *P1; g2m *Q0, *Q1; uint8_t *flp0, *flp1, *flq0, *flq1;
    const g1m *ng1; const uint8_t *flc;
};
// pks == nullptr: the public keys are already decoded and validated in b.Q1 / b.flq1 (min-sig) or b.P1 / b.flp1 (min-pk)
// `keys_sigs_ready`: the messages are on the device, the keys and signatures may still be on their way (host pipeline): hash
// first, wait for the event, then decode -- the upload of 192 of a tuple's 224 bytes hides under the hash kernel
int dev_verify_slice(Dev &d, const VerifyBufs &b, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *off, uint64_t base,
                     const uint8_t *sigs, size_t lo, size_t hi, const uint8_t *dst, uint32_t dst_len, int check, uint8_t *ok,
                     cudaEvent_t keys_sigs_ready = nullptr) {
    size_t n = hi - lo;
    const bool subgroup = (check & 2) != 0;
    check = (check & 1) | DEC_NO_IDENTITY;   // an identity key or signature never verifies (items.cuh: DEC_NO_IDENTITY)
    if (variant == BN254_MIN_SIG) {
        if (int rc = dev_hash_g1(d, b.P1 + lo, b.flp1 + lo, msgs, off + lo, base, n, dst, dst_len, 2)) return rc;
        if (keys_sigs_ready) CK(d, cudaStreamWaitEvent(d.cur, keys_sigs_ready, 0));
        LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, b.P0 + lo, b.flp0 + lo, sigs + 64 * lo, n, check);
        if (pks) {
            LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, b.Q1 + lo, b.flq1 + lo, pks + 128 * lo, n, check);
            if (subgroup) {
                LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, b.flq1 + lo, b.Q1 + lo, n);
            }
        }
        if (int rc = launch_miller2_fixed(d, n, b.F + lo, b.flf + lo, b.P0 + lo, b.flp0 + lo, b.P1 + lo, b.flp1 + lo, b.Q1 + lo, b.flq1 + lo)) return rc;
    } else {
        if (int rc = dev_hash_g2(d, b.Q1 + lo, b.flq1 + lo, msgs, off + lo, base, n, dst, dst_len, 2)) return rc;
        if (keys_sigs_ready) CK(d, cudaStreamWaitEvent(d.cur, keys_sigs_ready, 0));
        LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, b.Q0 + lo, b.flq0 + lo, sigs + 128 * lo, n, check);
        if (pks) LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, b.P1 + lo, b.flp1 + lo, pks + 64 * lo, n, check);
        if (subgroup) {
            LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, b.flq0 + lo, b.Q0 + lo, n);
        }
        if (int rc = launch_miller2(d, n, b.F + lo, b.flf + lo, b.ng1, b.flc, (size_t)0, b.Q0 + lo, b.flq0 + lo, b.P1 + lo, b.flp1 + lo, b.Q1 + lo,
                                    b.flq1 + lo))
            return rc;
    }
    if (int rc = launch_final_exp(d, b.F + lo, n)) return rc;
    LAUNCH(d, k_fp12_is_one, n, BN_TPB_LIGHT, ok + lo, b.F + lo, b.flf + lo, n);
    return 0;
}
// verification of one chunk, device pointers in / out
int dev_verify(Dev &d, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *off, uint64_t base, const uint8_t *sigs,
               size_t n, const uint8_t *dst, uint32_t dst_len, uint32_t flags, uint8_t *ok, cudaEvent_t keys_sigs_ready = nullptr) {
    int check = ((flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? 1 : 0) | ((flags & BN254_FLAG_CHECK_SUBGROUP) ? 2 : 0);
    VerifyBufs b = {};
    if (int rc = ensure_t(d, S_F, n, &b.F)) return rc;
    if (int rc = ensure_t(d, S_FLF, n, &b.flf)) return rc;
    if (int rc = ensure_t(d, S_P1, n, &b.P1)) return rc;
    if (int rc = ensure_t(d, S_Q1, n, &b.Q1)) return rc;
    if (int rc = ensure_t(d, S_FLP1, n, &b.flp1)) return rc;
    if (int rc = ensure_t(d, S_FLQ1, n, &b.flq1)) return rc;
    if (variant == BN254_MIN_SIG) {
        if (int rc = ensure_t(d, S_P0, n, &b.P0)) return rc;
        if (int rc = ensure_t(d, S_FLP0, n, &b.flp0)) return rc;
    } else {
        if (int rc = ensure_t(d, S_Q0, n, &b.Q0)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, n, &b.flq0)) return rc;
        if (int rc = neg_g1_const(d, &b.ng1, &b.flc)) return rc;
    }
    int S = d.nsub;
    if (d.prof || n < (size_t)2048 * (size_t)S) S = 1;   // per-kernel timing wants the kernels serial
    if (S <= 1) return dev_verify_slice(d, b, variant, pks, msgs, off, base, sigs, 0, n, dst, dst_len, check, ok, keys_sigs_ready);
    // fork: every slice waits for what is already queued on the main stream, then runs its own pipeline
    CK(d, cudaEventRecord(d.ev_fork, d.st));
    int rc = 0;
    for (int k = 0; k < S && !rc; k++) {
        size_t lo = n * (size_t)k / S, hi = n * (size_t)(k + 1) / S;
        d.cur = d.sub[k];
        cudaError_t e = cudaStreamWaitEvent(d.cur, d.ev_fork, 0);
        if (e != cudaSuccess) { d.cur = d.st; CK(d, e); }
        rc = dev_verify_slice(d, b, variant, pks, msgs, off, base, sigs, lo, hi, dst, dst_len, check, ok, keys_sigs_ready);
        if (!rc && cudaEventRecord(d.ev_join[k], d.cur) != cudaSuccess) rc = BN254_ERR_CUDA_BASE;
    }
    d.cur = d.st;
    if (rc) return rc;
    for (int k = 0; k < S; k++) CK(d, cudaStreamWaitEvent(d.st, d.ev_join[k], 0));   // join
    return 0;
}

// min-sig verification of one chunk against PREPARED public keys (G2Prepared): decode the signatures, hash, the two-pair loop
// with both line sequences precomputed (no twist-point arithmetic), final exponentiation, compare
int dev_verify_prepared(Dev &d, const line_t<fp2> *tab, const uint8_t *flk, const g2m *pts, const uint32_t *key_index, const uint8_t *msgs, const uint64_t *off,
                        uint64_t base, const uint8_t *sigs, size_t n, const uint8_t *dst, uint32_t dst_len, uint32_t flags, uint8_t *ok) {
    const int check = ((flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? DEC_ON_CURVE : 0) | DEC_NO_IDENTITY;
    fp12 *F; uint8_t *flf, *flp0, *flp1; g1m *P0, *P1;
    if (int rc = ensure_t(d, S_F, n, &F)) return rc;
    if (int rc = ensure_t(d, S_FLF, n, &flf)) return rc;
    if (int rc = ensure_t(d, S_P0, n, &P0)) return rc;
    if (int rc = ensure_t(d, S_FLP0, n, &flp0)) return rc;
    if (int rc = ensure_t(d, S_P1, n, &P1)) return rc;
    if (int rc = ensure_t(d, S_FLP1, n, &flp1)) return rc;
    LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P0, flp0, sigs, n, check);
    if (int rc = dev_hash_g1(d, P1, flp1, msgs, off, base, n, dst, dst_len, 2)) return rc;
    if (use_wide(d, n)) {
        // A batch this small is latency bound: the line tables save work, not time.  Fetch the keys of the tuples and run the
        // ordinary two-pair loop in its wide shape (the same bytes: the table holds exactly the lines that loop computes).
        g2m *Q1; uint8_t *flq1;
        if (int rc = ensure_t(d, S_Q1, n, &Q1)) return rc;
        if (int rc = ensure_t(d, S_FLQ1, n, &flq1)) return rc;
        LAUNCH(d, k_g2_fetch, n, BN_TPB_LIGHT, Q1, flq1, pts, flk, key_index, n);
        if (int rc = launch_miller2_fixed(d, n, F, flf, P0, flp0, P1, flp1, Q1, flq1)) return rc;
    } else {
        LAUNCH_PAIR(d, k_miller2_prep, k_miller2_prep_l2, n, F, flf, P0, flp0, P1, flp1, tab, flk, key_index, n);
    }
    if (int rc = launch_final_exp(d, F, n)) return rc;
    LAUNCH(d, k_fp12_is_one, n, BN_TPB_LIGHT, ok, F, flf, n);
    return 0;
}

// FastAggregateVerify for G groups (device pointers): decode + validate all keys, the aggregate key of every group (one thread per
// group), then the ordinary verification pipeline on (aggregate key, message, aggregate signature) with the keys pre-decoded
int dev_fast_aggregate(Dev &d, int variant, const uint8_t *pks, size_t nkeys, const uint64_t *koff, uint64_t kbase, const uint8_t *msgs,
                       const uint64_t *off, uint64_t base, const uint8_t *sigs, size_t G, const uint8_t *dst, uint32_t dst_len, uint32_t flags, uint8_t *ok) {
    const int check = ((flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? DEC_ON_CURVE : 0) | DEC_NO_IDENTITY;
    const bool subgroup = (flags & BN254_FLAG_CHECK_SUBGROUP) != 0;
    uint8_t *flk, *flsum;
    if (int rc = ensure_t(d, S_FLTMP, nkeys + 1, &flk)) return rc;
    if (variant == BN254_MIN_SIG) {
        g2m *K, *Q1;
        if (int rc = ensure_t(d, S_TMP, (nkeys + 1) * sizeof(g2m), (uint8_t **)&K)) return rc;
        if (int rc = ensure_t(d, S_Q1, G, &Q1)) return rc;
        if (int rc = ensure_t(d, S_FLQ1, G, &flsum)) return rc;
        LAUNCH(d, k_g2_decode, nkeys, BN_TPB_LIGHT, K, flk, pks, nkeys, check);
        if (subgroup) LAUNCH(d, k_g2_require_subgroup, nkeys, BN_TPB_HEAVY, flk, K, nkeys);
        LAUNCH(d, k_g2_seg_sum_var, G, BN_TPB_HEAVY, Q1, flsum, K, flk, koff, kbase, G);
    } else {
        g1m *K, *P1;
        if (int rc = ensure_t(d, S_TMP, (nkeys + 1) * sizeof(g1m), (uint8_t **)&K)) return rc;
        if (int rc = ensure_t(d, S_P1, G, &P1)) return rc;
        if (int rc = ensure_t(d, S_FLP1, G, &flsum)) return rc;
        LAUNCH(d, k_g1_decode, nkeys, BN_TPB_LIGHT, K, flk, pks, nkeys, check);
        LAUNCH(d, k_g1_seg_sum_var, G, BN_TPB_HEAVY, P1, flsum, K, flk, koff, kbase, G);
    }
    return dev_verify(d, variant, nullptr, msgs, off, base, sigs, G, dst, dst_len, flags, ok);
}

// ---- randomised batch verification ----------------------------------------------
// Segmented reduction of n records in groups of `group` (the last group may be short) down to one record per
// group: levels of fan-in BN_SEG_FAN (4: the tail is latency-bound, 16-way levels measured 0.8 ms slower), ping-pong between A and B.  *res = nseg contiguous records.
#ifndef BN_RBV_TWIN
#define BN_RBV_TWIN 2
#endif
#ifndef BN_SEG_FAN
#define BN_SEG_FAN 4
#endif
#define SEG_LEVELS(d, kern, T, in, A, B, n, group, res)                                                   \
    do {                                                                                                  \
        seg_level L_;                                                                                     \
        L_.nseg = ((n) + (group) - 1) / (group);                                                          \
        L_.stride_in = (group);                                                                           \
        L_.len_full = (group);                                                                            \
        L_.len_last = (n) - (L_.nseg - 1) * (group);                                                      \
        L_.fan = BN_SEG_FAN;                                                                              \
        const T *src_ = (in);                                                                             \
        T *dst_ = (A);                                                                                    \
        while (L_.len_full > 1) {                                                                         \
            L_.stride_out = (L_.len_full + BN_SEG_FAN - 1) / BN_SEG_FAN;                                  \
            LAUNCH(d, kern, L_.nseg * L_.stride_out, BN_TPB_HEAVY, dst_, src_, L_);                       \
            src_ = dst_;                                                                                  \
            dst_ = (dst_ == (A)) ? (B) : (A);                                                             \
            L_.stride_in = L_.stride_out;                                                                 \
            L_.len_full = L_.stride_out;                                                                  \
            L_.len_last = (L_.len_last + BN_SEG_FAN - 1) / BN_SEG_FAN;                                    \
        }                                                                                                 \
        *(res) = const_cast<T *>(src_);                                                                   \
    } while (0)

// One chunk: n tuples (device pointers) in groups of `group` -> ceil(n / group) accept bytes.  `idx0` is the
// index of the chunk's first tuple in the whole call (it enters the scalar derivation).
// `tail_set` >= 0 (the host pipeline): the per-group tail -- final exponentiation and verdict -- runs on d.tail with buffers of
// that staging set, so that it overlaps the main kernels of the next chunk; the caller collects `ok` on d.tail.
int dev_verify_rand(Dev &d, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *off, uint64_t base, const uint8_t *sigs,
                    size_t n, size_t group, const uint8_t *seed, uint64_t idx0, int nbits, const uint8_t *dst, uint32_t dst_len, uint32_t flags,
                    uint8_t *ok, int tail_set = -1, bool tail_hidden = false) {
    const int check = ((flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? DEC_ON_CURVE : 0) | DEC_NO_IDENTITY;
    const bool subgroup = (flags & BN254_FLAG_CHECK_SUBGROUP) != 0;
    const size_t nseg = (n + group - 1) / group;
    const size_t so0 = (group + BN_SEG_FAN - 1) / BN_SEG_FAN, so1 = (so0 + BN_SEG_FAN - 1) / BN_SEG_FAN;
    fp12 *F, *FA, *FB, *Fs, *Fg;
    g1m *P1, *GP; g2m *Q1, *GQ;
    uint8_t *flp1, *flq1, *gflp, *gflq, *flsig;
    int *gbad;
    if (int rc = ensure_t(d, S_F, n, &F)) return rc;
    if (int rc = ensure_t(d, S_RB_FA, nseg * so0 + 1, &FA)) return rc;
    if (int rc = ensure_t(d, S_RB_FB, nseg * so1 + 1, &FB)) return rc;
    if (int rc = ensure_t(d, S_RB_FS, nseg, &Fs)) return rc;
    if (int rc = ensure_t(d, S_P1, n, &P1)) return rc;
    if (int rc = ensure_t(d, S_Q1, n, &Q1)) return rc;
    if (int rc = ensure_t(d, S_FLP1, n, &flp1)) return rc;
    if (int rc = ensure_t(d, S_FLQ1, n, &flq1)) return rc;
    if (int rc = ensure_t(d, S_RB_GP, nseg, &GP)) return rc;
    if (int rc = ensure_t(d, S_RB_GQ, nseg, &GQ)) return rc;
    if (int rc = ensure_t(d, S_RB_FLP, nseg, &gflp)) return rc;
    if (int rc = ensure_t(d, S_RB_FLQ, nseg, &gflq)) return rc;
    if (tail_set >= 0 && (d.prof || d.shape == 3 || d.cur != d.st)) tail_set = -1;   // per-kernel timing / shared scratch want it serial
    // A tail that overlaps the next chunk's kernels is off the critical path: it keeps the throughput shape (the wide machine
    // buys latency with shared memory and registers the overlapped kernels want).
    struct WideOff {
        Dev &d; size_t keep;
        WideOff(Dev &dd, bool off) : d(dd), keep(dd.wide_max) { if (off) d.wide_max = 0; }
        ~WideOff() { d.wide_max = keep; }
    } wide_off(d, tail_set >= 0 && tail_hidden);
    if (int rc = ensure_t(d, stage_buf(S_RB_BAD, tail_set > 0 ? 1 : 0), nseg, &gbad)) return rc;
    CK(d, cudaMemsetAsync(gbad, 0, nseg * sizeof(int), d.cur));
    // The signature side (scale, per-group sums, one Miller loop per GROUP against the fixed generator) only needs
    // the decoded signatures: it runs on the side stream next to the per-tuple pipeline of the main stream.
    struct SideStream {   // LAUNCH goes to d.cur; always put the main stream back
        Dev &d; cudaStream_t keep;
        explicit SideStream(Dev &dd) : d(dd), keep(dd.cur) {}
        ~SideStream() { d.cur = keep; }
    };
    const bool side = !d.prof && d.cur == d.st;   // per-kernel timing wants the kernels serial
    if (variant == BN254_MIN_SIG) {
        g1m *P0; g1p *S0, *SA, *SB, *Sg;
        if (int rc = ensure_t(d, S_P0, n, &P0)) return rc;
        if (int rc = ensure_t(d, S_FLP0, n, &flsig)) return rc;
        if (int rc = ensure_t(d, S_RB_S0, n, &S0)) return rc;
        if (int rc = ensure_t(d, S_RB_S1, nseg * so0 + 1, &SA)) return rc;
        if (int rc = ensure_t(d, S_RB_S2, nseg * so1 + 1, &SB)) return rc;
        LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P0, flsig, sigs, n, check);
        {
            SideStream guard(d);
            if (side) {
                CK(d, cudaEventRecord(d.ev_hi_fork, d.st));
                CK(d, cudaStreamWaitEvent(d.hi, d.ev_hi_fork, 0));
                d.cur = d.hi;
            }
            LAUNCH(d, k_rbv_scale_g1p, n, BN_TPB_HEAVY, S0, P0, flsig, seed, idx0, nbits, n);
            SEG_LEVELS(d, k_g1p_seg_sum, g1p, S0, SA, SB, n, group, &Sg);
            LAUNCH(d, k_g1p_to_affine, nseg, BN_TPB_HEAVY, GP, gflp, Sg, nseg);
            LAUNCH(d, k_fill_neg_g2, nseg, BN_TPB_LIGHT, GQ, gflq, nseg);
            {   // runs beside the per-tuple pipeline of the main stream: throughput shape (measured: a wide group loop costs the
                // main stream more than it saves, 19.6 -> 21.0 ms per 2^16 tuples)
                WideOff beside(d, true);
                if (int rc = launch_miller1(d, nseg, Fs, (uint8_t *)nullptr, GP, gflp, GQ, gflq, (int *)nullptr)) return rc;
            }
            if (side) CK(d, cudaEventRecord(d.ev_hi_join, d.hi));
        }
        LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, Q1, flq1, pks, n, check);
        if (subgroup) LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, flq1, Q1, n);
        if (int rc = dev_hash_g1(d, P1, flp1, msgs, off, base, n, dst, dst_len, 2)) return rc;
    } else {
        g2m *Q0; g2p *S0, *SA, *SB, *Sg;
        if (int rc = ensure_t(d, S_Q0, n, &Q0)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, n, &flsig)) return rc;
        if (int rc = ensure_t(d, S_RB_S0, n, &S0)) return rc;
        if (int rc = ensure_t(d, S_RB_S1, nseg * so0 + 1, &SA)) return rc;
        if (int rc = ensure_t(d, S_RB_S2, nseg * so1 + 1, &SB)) return rc;
        LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, Q0, flsig, sigs, n, check);
        if (subgroup) LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, flsig, Q0, n);
        {
            SideStream guard(d);
            if (side) {
                CK(d, cudaEventRecord(d.ev_hi_fork, d.st));
                CK(d, cudaStreamWaitEvent(d.hi, d.ev_hi_fork, 0));
                d.cur = d.hi;
            }
            LAUNCH(d, k_rbv_scale_g2p, n, BN_TPB_HEAVY, S0, Q0, flsig, seed, idx0, nbits, n);
            SEG_LEVELS(d, k_g2p_seg_sum, g2p, S0, SA, SB, n, group, &Sg);
            LAUNCH(d, k_g2p_to_affine, nseg, BN_TPB_HEAVY, GQ, gflq, Sg, nseg);
            LAUNCH(d, k_fill_neg_g1, nseg, BN_TPB_LIGHT, GP, gflp, nseg);
            {   // runs beside the per-tuple pipeline of the main stream: throughput shape (measured: a wide group loop costs the
                // main stream more than it saves, 19.6 -> 21.0 ms per 2^16 tuples)
                WideOff beside(d, true);
                if (int rc = launch_miller1(d, nseg, Fs, (uint8_t *)nullptr, GP, gflp, GQ, gflq, (int *)nullptr)) return rc;
            }
            if (side) CK(d, cudaEventRecord(d.ev_hi_join, d.hi));
        }
        LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P1, flp1, pks, n, check);
        if (int rc = dev_hash_g2(d, Q1, flq1, msgs, off, base, n, dst, dst_len, 2)) return rc;
    }
    LAUNCH(d, k_rbv_group_flags, n, BN_TPB_LIGHT, gbad, flsig, flp1, flq1, group, n);
    LAUNCH(d, k_rbv_scale_g1, n, BN_TPB_HEAVY, P1, flp1, seed, idx0, nbits, n);
    // Even groups: two consecutive tuples share one two-pair Miller loop (one Fp12 squaring per step for both), which
    // is also the first level of the group product.  Measured on B200: +10 % at 2^18 tuples, +11.5 % at 2^20, but -3.5 % at
    // 2^16 and worse below (half as many threads for a machine that is not full).  BN_RBV_TWIN: 0 never, 1 whenever the
    // group size is even, 2 (default) when, in addition, one thread per tuple would overfill one wave of lane pairs.
    if (group % 2 == 0 && (BN_RBV_TWIN == 1 || (BN_RBV_TWIN == 2 && 2 * n > d.lane_wave_threads))) {
        const size_t half = (n + 1) / 2;
        if (int rc = launch_miller_twin(d, n, F, P1, flp1, Q1, flq1, (int *)nullptr)) return rc;
        SEG_LEVELS(d, k_fp12_seg_reduce, fp12, F, FA, FB, half, group / 2, &Fg);
    } else {
        if (int rc = launch_miller1(d, n, F, (uint8_t *)nullptr, P1, flp1, Q1, flq1, (int *)nullptr)) return rc;
        SEG_LEVELS(d, k_fp12_seg_reduce, fp12, F, FA, FB, n, group, &Fg);
    }
    if (side) CK(d, cudaStreamWaitEvent(d.st, d.ev_hi_join, 0));
    if (tail_set < 0) {
        LAUNCH(d, k_fp12_mul_into, nseg, BN_TPB_HEAVY, Fg, Fs, nseg);
        if (int rc = launch_final_exp(d, Fg, nseg)) return rc;
        LAUNCH(d, k_rbv_ok, nseg, BN_TPB_LIGHT, ok, Fg, gbad, nseg);
        return 0;
    }
    // hand the group products to the tail stream in a buffer of this staging set; everything the next chunk overwrites
    // (F, FA, FB, Fs) has been consumed on the launch stream by then
    fp12 *T;
    if (int rc = ensure_t(d, stage_buf(S_RB_T, tail_set), nseg, &T)) return rc;
    LAUNCH(d, k_fp12_mul_out, nseg, BN_TPB_HEAVY, T, Fg, Fs, nseg);
    CK(d, cudaEventRecord(d.ev_tail[tail_set], d.st));
    CK(d, cudaStreamWaitEvent(d.tail, d.ev_tail[tail_set], 0));
    struct Cur {   // LAUNCH goes to d.cur; always put the main stream back
        Dev &d; cudaStream_t keep;
        explicit Cur(Dev &dd) : d(dd), keep(dd.cur) { dd.cur = dd.tail; }
        ~Cur() { d.cur = keep; }
    } guard(d);
    if (int rc = launch_final_exp(d, T, nseg)) return rc;
    LAUNCH(d, k_rbv_ok, nseg, BN_TPB_LIGHT, ok, T, gbad, nseg);
    return 0;
}

// ---- same-key randomised verification ("same-key grouping") ------------------------------------------
// Reduction plan of one chunk, computed on the host from the key offsets: a blob of u64 arrays -- the nk + 1 key offsets relative
// to the chunk, then per level the nb + 1 bounds of its blocks (fan-in BN_SEG_FAN, a block never straddles two keys, a key
// without tuples has no block), last the nk + 1 offsets of the keys' final records.
struct SkPlan {
    struct Level { size_t pos, nb; };
    std::vector<uint64_t> blob;
    std::vector<Level> levels;
    size_t kfin_pos = 0, nfin = 0, nb_max[2] = {0, 0};   // nb_max[j]: most blocks among the levels written to ping-pong buffer j
};
// a chunk of n tuples under nk keys takes the bucket path (no levels needed: the bucket kernels read the key offsets)
inline bool sk_use_msm(const Dev &d, size_t n, size_t nk) { return d.sk_msm_min && n >= d.sk_msm_min * nk; }
void sk_plan(SkPlan &p, const uint64_t *koff, size_t nk, bool with_levels) {
    const uint64_t t0 = koff[0];
    std::vector<uint64_t> cur(nk + 1), nxt(nk + 1);
    for (size_t k = 0; k <= nk; k++) cur[k] = koff[k] - t0;
    p.blob.assign(cur.begin(), cur.end());
    while (with_levels) {
        uint64_t longest = 0;
        for (size_t k = 0; k < nk; k++) longest = std::max(longest, cur[k + 1] - cur[k]);
        if (longest <= 1) break;
        SkPlan::Level L;
        L.pos = p.blob.size();
        nxt[0] = 0;
        for (size_t k = 0; k < nk; k++) {
            for (uint64_t s = cur[k]; s < cur[k + 1]; s += BN_SEG_FAN) p.blob.push_back(s);
            nxt[k + 1] = nxt[k] + (cur[k + 1] - cur[k] + BN_SEG_FAN - 1) / BN_SEG_FAN;
        }
        p.blob.push_back(cur[nk]);
        L.nb = (size_t)nxt[nk];
        size_t &m = p.nb_max[p.levels.size() & 1];
        m = std::max(m, L.nb);
        p.levels.push_back(L);
        cur.swap(nxt);
    }
    p.kfin_pos = p.blob.size();
    p.blob.insert(p.blob.end(), cur.begin(), cur.end());
    p.nfin = (size_t)cur[nk];
}
// One chunk of whole keys (device pointers; `plan` = the uploaded blob of `hp`): nk keys, their n tuples -> nk accept bytes.
//   min-sig: e(sum r_i sig_i, -g2) * e(sum r_i H(m_i), pk_k) == 1        min-pk: e(-g1, sum r_i sig_i) * e(pk_k, sum r_i H(m_i)) == 1
// Per tuple: decode, hash, two GLV scalings; per key: the two sums, ONE two-pair Miller loop and ONE final exponentiation --
// the exact predicate's own kernels on (pk_k, sum r_i H(m_i), sum r_i sig_i).  idx0 = index of the chunk's first tuple in the call.
// `tail_set` >= 0 (the host pipeline): the per-key tail -- Miller loop, final exponentiation, verdict: a latency-bound chain on
// a few keys -- runs on d.tail over buffers of that staging set and overlaps the per-tuple kernels of the next chunk; the caller
// collects `ok` on d.tail.
int dev_verify_same_key(Dev &d, int variant, const uint8_t *pks, size_t nk, const uint64_t *plan, const SkPlan &hp, const uint8_t *msgs,
                        const uint64_t *off, uint64_t base, const uint8_t *sigs, size_t n, const uint8_t *seed, uint64_t idx0, int nbits,
                        const uint8_t *dst, uint32_t dst_len, uint32_t flags, uint8_t *ok, int tail_set = -1, bool tail_hidden = false) {
    const int check = ((flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? DEC_ON_CURVE : 0) | DEC_NO_IDENTITY;
    const bool subgroup = (flags & BN254_FLAG_CHECK_SUBGROUP) != 0;
    const bool sig_g1 = variant == BN254_MIN_SIG;
    if (tail_set >= 0 && (d.prof || d.shape == 3 || d.cur != d.st)) tail_set = -1;   // per-kernel timing / shared scratch want it serial
    // everything the tail reads lives in ONE per-staging-set allocation: F | sums of the signature side | sums of the hash side |
    // decoded keys | four flag arrays
    const size_t nkp = (nk + 15) & ~(size_t)15;
    const size_t rec_sum = sig_g1 ? sizeof(g1m) : sizeof(g2m), rec_pk = sig_g1 ? sizeof(g2m) : sizeof(g1m);
    uint8_t *tb;
    if (int rc = ensure_t(d, stage_buf(S_RB_T, tail_set > 0 ? 1 : 0), nkp * (sizeof(fp12) + 2 * rec_sum + rec_pk + 4), &tb)) return rc;
    fp12 *F = (fp12 *)tb;
    uint8_t *ka = tb + nkp * sizeof(fp12), *kb = ka + nkp * rec_sum, *pk = kb + nkp * rec_sum;
    uint8_t *flf = pk + nkp * rec_pk, *flka = flf + nkp, *flkb = flka + nkp, *flpk = flkb + nkp;
    uint8_t *fla, *flb;
    int *kbad;
    if (int rc = ensure_t(d, S_RB_BAD, nk, &kbad)) return rc;
    CK(d, cudaMemsetAsync(kbad, 0, nk * sizeof(int), d.cur));
    const uint64_t *ko = plan, *kfin = plan + hp.kfin_pos;
    const g1m *ng1 = nullptr; const uint8_t *flc = nullptr;
    // Long segments: the two sums of a key as bucket-method multi-scalar multiplications (kernels.cuh: k_sk_bucket_*), ~4 x less
    // per-tuple work than scaling every point; short segments: scale every point, then the planned segmented sums.
    // "sk_msm_min" = smallest average segment length that takes the bucket path (0: never).
    const int NW = nbits >> 3;                                     // radix-16 windows per scalar half
    const bool msm = sk_use_msm(d, n, nk);
    const size_t nwin = 2 * nk * 2 * (size_t)NW;                   // (array, key, half, window)
    int msm_S = 1;                                                 // slices per segment: enough bucket threads to fill the machine
    if (msm) {
        const size_t avg = n / nk, want = (d.sms * 1024 + nwin * 16 - 1) / (nwin * 16);
        msm_S = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(want, avg / 64), 8));
        if (d.sk_msm_slices) msm_S = (int)d.sk_msm_slices;
    }
    uint8_t *D = nullptr;                                          // digit bytes: 2 NW arrays of npad
    const size_t npad = (n + 15) & ~(size_t)15;
    if (msm)
        if (int rc = ensure_t(d, S_F2, 2 * (size_t)NW * npad + 16, &D)) return rc;
    if (sig_g1) {
        g1m *A, *B; g1p *S, *SA, *SB;
        if (int rc = ensure_t(d, S_P0, n, &A)) return rc;
        if (int rc = ensure_t(d, S_FLP0, n, &fla)) return rc;
        if (int rc = ensure_t(d, S_P1, n, &B)) return rc;
        if (int rc = ensure_t(d, S_FLP1, n, &flb)) return rc;
        if (int rc = ensure_t(d, S_RB_S0, msm ? nwin * 16 * msm_S : 2 * n + 1, &S)) return rc;
        if (int rc = ensure_t(d, S_RB_S1, msm ? nwin : 2 * hp.nb_max[0] + 1, &SA)) return rc;
        if (int rc = ensure_t(d, S_RB_S2, msm ? 2 * nk : 2 * hp.nb_max[1] + 1, &SB)) return rc;
        LAUNCH(d, k_g2_decode, nk, BN_TPB_LIGHT, (g2m *)pk, flpk, pks, nk, check);
        if (subgroup) LAUNCH(d, k_g2_require_subgroup, nk, BN_TPB_HEAVY, flpk, (g2m *)pk, nk);
        const g1p *src = S;
        if (n) {
            LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, A, fla, sigs, n, check);
            if (int rc = dev_hash_g1(d, B, flb, msgs, off, base, n, dst, dst_len, 2)) return rc;
            LAUNCH(d, k_sk_key_flags, n, BN_TPB_LIGHT, kbad, fla, flb, ko, nk, n);
            if (msm) {
                LAUNCH(d, k_sk_digits, n, BN_TPB_LIGHT, D, seed, idx0, nbits, n, npad);
                LAUNCH(d, k_sk_bucket_g1, nwin * 16 * msm_S, BN_TPB_HEAVY, S, A, fla, B, flb, D, npad, ko, nk, NW, msm_S);
                LAUNCH(d, k_sk_window_g1, (nwin * 16 + 31) & ~(size_t)31, BN_TPB_HEAVY, SA, S, nwin, msm_S);
                LAUNCH(d, k_sk_combine_g1, 2 * nk, BN_TPB_HEAVY, SB, SA, nk, NW);
                src = SB;
            } else {
                LAUNCH(d, k_rbv_scale_g1p, n, BN_TPB_HEAVY, S, A, fla, seed, idx0, nbits, n);
                LAUNCH(d, k_rbv_scale_g1p, n, BN_TPB_HEAVY, S + n, B, flb, seed, idx0, nbits, n);
                size_t n_in = n;
                for (size_t l = 0; l < hp.levels.size(); l++) {
                    g1p *out = (l & 1) ? SB : SA;
                    LAUNCH(d, k_sk_sum_g1p, 2 * hp.levels[l].nb, BN_TPB_HEAVY, out, src, plan + hp.levels[l].pos, hp.levels[l].nb, n_in);
                    src = out;
                    n_in = hp.levels[l].nb;
                }
            }
        }
        LAUNCH(d, k_sk_finish_g1, nk, BN_TPB_HEAVY, (g1m *)ka, flka, (g1m *)kb, flkb, src, msm ? ko : kfin, kbad, nk, msm);
    } else {
        g2m *A, *B; g2p *S, *SA, *SB;
        if (int rc = ensure_t(d, S_Q0, n, &A)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, n, &fla)) return rc;
        if (int rc = ensure_t(d, S_Q1, n, &B)) return rc;
        if (int rc = ensure_t(d, S_FLQ1, n, &flb)) return rc;
        if (int rc = ensure_t(d, S_RB_S0, msm ? nwin * 16 * msm_S : 2 * n + 1, &S)) return rc;
        if (int rc = ensure_t(d, S_RB_S1, msm ? nwin : 2 * hp.nb_max[0] + 1, &SA)) return rc;
        if (int rc = ensure_t(d, S_RB_S2, msm ? 2 * nk : 2 * hp.nb_max[1] + 1, &SB)) return rc;
        if (int rc = neg_g1_const(d, &ng1, &flc)) return rc;
        LAUNCH(d, k_g1_decode, nk, BN_TPB_LIGHT, (g1m *)pk, flpk, pks, nk, check);
        const g2p *src = S;
        if (n) {
            LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, A, fla, sigs, n, check);
            if (subgroup) LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, fla, A, n);
            if (int rc = dev_hash_g2(d, B, flb, msgs, off, base, n, dst, dst_len, 2)) return rc;
            LAUNCH(d, k_sk_key_flags, n, BN_TPB_LIGHT, kbad, fla, flb, ko, nk, n);
            if (msm) {
                LAUNCH(d, k_sk_digits, n, BN_TPB_LIGHT, D, seed, idx0, nbits, n, npad);
                LAUNCH(d, k_sk_bucket_g2, nwin * 16 * msm_S, BN_TPB_HEAVY, S, A, fla, B, flb, D, npad, ko, nk, NW, msm_S);
                LAUNCH(d, k_sk_window_g2, (nwin * 16 + 31) & ~(size_t)31, BN_TPB_HEAVY, SA, S, nwin, msm_S);
                LAUNCH(d, k_sk_combine_g2, 2 * nk, BN_TPB_HEAVY, SB, SA, nk, NW);
                src = SB;
            } else {
                LAUNCH(d, k_rbv_scale_g2p, n, BN_TPB_HEAVY, S, A, fla, seed, idx0, nbits, n);
                LAUNCH(d, k_rbv_scale_g2p, n, BN_TPB_HEAVY, S + n, B, flb, seed, idx0, nbits, n);
                size_t n_in = n;
                for (size_t l = 0; l < hp.levels.size(); l++) {
                    g2p *out = (l & 1) ? SB : SA;
                    LAUNCH(d, k_sk_sum_g2p, 2 * hp.levels[l].nb, BN_TPB_HEAVY, out, src, plan + hp.levels[l].pos, hp.levels[l].nb, n_in);
                    src = out;
                    n_in = hp.levels[l].nb;
                }
            }
        }
        LAUNCH(d, k_sk_finish_g2, nk, BN_TPB_HEAVY, (g2m *)ka, flka, (g2m *)kb, flkb, src, msm ? ko : kfin, kbad, nk, msm);
    }
    // ---- the per-key tail: on the launch stream, or on d.tail beside the next chunk.  tail_hidden would put a hidden tail into
    //      the throughput shape (as the randomised path does for its per-group tail); here the host entry leaves it off: a few
    //      dozen keys per chunk are a few dozen warps in the wide shape, and its 3.3 ms latency (lane pairs: 8 ms, as long as a
    //      whole chunk) lets staging set k be handed back before chunk k + 1 ends, so the host can queue chunk k + 2 in time ----
    struct WideOff {
        Dev &d; size_t keep;
        WideOff(Dev &dd, bool off) : d(dd), keep(dd.wide_max) { if (off) d.wide_max = 0; }
        ~WideOff() { d.wide_max = keep; }
    } wide_off(d, tail_set >= 0 && tail_hidden);
    struct Cur {   // LAUNCH goes to d.cur; always put the main stream back
        Dev &d; cudaStream_t keep;
        Cur(Dev &dd, bool tail) : d(dd), keep(dd.cur) { if (tail) dd.cur = dd.tail; }
        ~Cur() { d.cur = keep; }
    };
    if (tail_set >= 0) {
        CK(d, cudaEventRecord(d.ev_tail[tail_set], d.st));
        CK(d, cudaStreamWaitEvent(d.tail, d.ev_tail[tail_set], 0));
    }
    Cur guard(d, tail_set >= 0);
    if (sig_g1) {
        if (int rc = launch_miller2_fixed(d, nk, F, flf, (const g1m *)ka, flka, (const g1m *)kb, flkb, (const g2m *)pk, flpk)) return rc;
    } else {
        if (int rc = launch_miller2(d, nk, F, flf, ng1, flc, (size_t)0, (const g2m *)ka, flka, (const g1m *)pk, flpk, (const g2m *)kb, flkb)) return rc;
    }
    if (int rc = launch_final_exp(d, F, nk)) return rc;
    LAUNCH(d, k_fp12_is_one, nk, BN_TPB_LIGHT, ok, F, flf, nk);
    return 0;
}

// One shard of an aggregate verification (device pointers): rec->f = prod_i ML(H(m_i), pk_i) [* ML(aggsig, -g2) when
// agg_sig != NULL] in internal form, rec->bad = a key was undecodable / the identity / failed the requested checks, or a
// message hit the SvdW exception.  The last pass of the product tree writes rec->f in place (rec may be the NCCL send
// record).  flags: CHECK_SUBGROUP implies CHECK_ON_CURVE, as in verify; G1 has cofactor 1, so only G2 inputs (min-sig
// keys, a min-pk aggregate signature) need the psi criterion.
int dev_aggregate_rec(Dev &d, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *off, uint64_t base, size_t n,
                      const uint8_t *agg_sig, const uint8_t *dst, uint32_t dst_len, uint32_t flags, agg_rec *rec) {
    const int check = (flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? DEC_ON_CURVE : 0;
    const bool subgroup = (flags & BN254_FLAG_CHECK_SUBGROUP) != 0;
    size_t tot = n + (agg_sig ? 1 : 0);
    fp12 *F; g1m *P; g2m *Q; uint8_t *flp, *flq; int *dbad;
    if (int rc = ensure_t(d, S_F, tot + 1, &F)) return rc;
    if (int rc = ensure_t(d, S_P1, tot + 1, &P)) return rc;
    if (int rc = ensure_t(d, S_Q1, tot + 1, &Q)) return rc;
    if (int rc = ensure_t(d, S_FLP1, tot + 1, &flp)) return rc;
    if (int rc = ensure_t(d, S_FLQ1, tot + 1, &flq)) return rc;
    if (int rc = ensure_t(d, S_BAD, 4, &dbad)) return rc;
    CK(d, cudaMemsetAsync(dbad, 0, sizeof(int), d.cur));
    if (variant == BN254_MIN_SIG) {
        LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, Q, flq, pks, n, check | DEC_NO_IDENTITY);
        if (subgroup) LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, flq, Q, n);
        if (int rc = dev_hash_g1(d, P, flp, msgs, off, base, n, dst, dst_len, 2)) return rc;
        if (agg_sig) {
            LAUNCH(d, k_g1_decode, 1, BN_TPB_LIGHT, P + n, flp + n, agg_sig, (size_t)1, check);
            LAUNCH(d, k_fill_neg_g2, 1, BN_TPB_LIGHT, Q + n, flq + n, (size_t)1);
        }
    } else {
        LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P, flp, pks, n, check | DEC_NO_IDENTITY);
        if (int rc = dev_hash_g2(d, Q, flq, msgs, off, base, n, dst, dst_len, 2)) return rc;
        if (agg_sig) {
            LAUNCH(d, k_g2_decode, 1, BN_TPB_LIGHT, Q + n, flq + n, agg_sig, (size_t)1, check);
            if (subgroup) LAUNCH(d, k_g2_require_subgroup, 1, BN_TPB_HEAVY, flq + n, Q + n, (size_t)1);
            LAUNCH(d, k_fill_neg_g1, 1, BN_TPB_LIGHT, P + n, flp + n, (size_t)1);
        }
    }
    // a batch that overfills the machine takes two pairs per thread: one shared Fp12 squaring per step (see dev_verify_rand)
    size_t nf = tot;
    if (BN_RBV_TWIN == 1 || (BN_RBV_TWIN == 2 && 2 * tot > d.lane_wave_threads)) {
        nf = (tot + 1) / 2;
        if (int rc = launch_miller_twin(d, tot, F, P, flp, Q, flq, dbad)) return rc;
    } else {
        if (int rc = launch_miller1(d, tot, F, (uint8_t *)nullptr, P, flp, Q, flq, dbad)) return rc;
    }
    fp12 *res;
    if (int rc = dev_fp12_product(d, F, nf, &res, &rec->f)) return rc;
    LAUNCH(d, k_agg_rec_flag, 1, 32, rec, dbad);
    return 0;
}
// the same with the partial in wire form (384 bytes) and the flag as one byte: bn254_aggregate_partial[_dev]
int dev_aggregate_partial(Dev &d, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *off, uint64_t base, size_t n,
                          const uint8_t *agg_sig, const uint8_t *dst, uint32_t dst_len, uint32_t flags, uint8_t *partial, uint8_t *bad) {
    agg_rec *rec;
    if (int rc = ensure_t(d, S_AGG_SEND, 1, &rec)) return rc;
    if (int rc = dev_aggregate_rec(d, variant, pks, msgs, off, base, n, agg_sig, dst, dst_len, flags, rec)) return rc;
    LAUNCH(d, k_fp12_encode, 1, BN_TPB_LIGHT, partial, (uint8_t *)nullptr, &rec->f, (const uint8_t *)nullptr, (size_t)1);
    CK(d, cudaMemcpyAsync(bad, &rec->bad, 1, cudaMemcpyDeviceToDevice, d.cur));  // little-endian low byte
    return 0;
}

// partials (wire, device) -> ok byte (device)
int dev_aggregate_finish(Dev &d, const uint8_t *partials, const uint8_t *bad, size_t kparts, uint8_t *ok) {
    fp12 *F; uint8_t *flf; int *dbad;
    if (int rc = ensure_t(d, S_F, kparts + 1, &F)) return rc;
    if (int rc = ensure_t(d, S_FLF, kparts + 1, &flf)) return rc;
    if (int rc = ensure_t(d, S_BAD, 4, &dbad)) return rc;
    CK(d, cudaMemsetAsync(dbad, 0, sizeof(int), d.cur));
    LAUNCH(d, k_fp12_decode, kparts, BN_TPB_LIGHT, F, flf, partials, kparts);
    LAUNCH(d, k_or_flags, kparts, BN_TPB_LIGHT, dbad, flf, kparts);
    LAUNCH(d, k_or_bytes, kparts, BN_TPB_LIGHT, dbad, bad, kparts);
    fp12 *res;
    if (int rc = dev_fp12_product(d, F, kparts, &res)) return rc;
    if (int rc = launch_final_exp(d, res, (size_t)1)) return rc;
    LAUNCH(d, k_finish_ok, 1, 32, ok, res, dbad);
    return 0;
}

// ---- the aggregate path's exchange: ONE ncclAllGather of 400 bytes per rank on the launch stream, then
//      ONE kernel (product of <= nranks partials, the single final exponentiation, compare).  No host round trip.
#define NCCLCK(d, call)                                                                                  \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != bn_nccl::kSuccess) {                                                                   \
            (d).err = std::string(#call) + ": " + (bn_nccl::api().GetErrorString ? bn_nccl::api().GetErrorString(r_) : "NCCL error"); \
            (d).rc = BN254_ERR_NCCL_BASE - r_;                                                           \
            return (d).rc;                                                                               \
        }                                                                                                \
    } while (0)
int need_comm(Dev &d) {
    if (d.comm) return 0;
    d.err = bn_nccl::api().ok() ? "no communicator: call bn254_comm_init_rank (or create the ctx with several devices)"
                                : (bn_nccl::api().why ? bn_nccl::api().why : "NCCL unavailable");
    d.rc = BN254_ERR_NCCL_BASE;
    return d.rc;
}
// this rank's record (device, S_AGG_SEND) -> every rank's records (device, S_AGG_RECV); enqueued on d.cur
int dev_gather_recs(Dev &d, const agg_rec *send, agg_rec **recv) {
    if (int rc = ensure_t(d, S_AGG_RECV, (size_t)d.nranks, recv)) return rc;
    if (d.nranks == 1) {
        CK(d, cudaMemcpyAsync(*recv, send, sizeof(agg_rec), cudaMemcpyDeviceToDevice, d.cur));
        return 0;
    }
    if (int rc = need_comm(d)) return rc;
    NCCLCK(d, bn_nccl::api().AllGather(send, *recv, sizeof(agg_rec), bn_nccl::kUint8, d.comm, d.cur));
    d.launches++;   // NCCL's kernel
    return 0;
}
int dev_gather_finish(Dev &d, const agg_rec *recs, uint8_t *ok) {
    if (use_wide(d, 1)) {
        LAUNCH_W(d, k_w_gather_finish, R_FE_CYCSQR, R_FE_STF, 1, ok, (fp12 *)nullptr, recs, d.nranks);
        return 0;
    }
    LAUNCH(d, k_gather_finish_l2, 32, 32, ok, (fp12 *)nullptr, recs, d.nranks);
    return 0;
}

// ---- sharding helpers -----------------------------------------------------------------------------
template <class Fn>
int run_sharded(bn254_ctx *ctx, size_t n, Fn fn) {
    size_t nd = ctx->devs.size();
    if (nd == 1 || n < nd) {
        Dev &d = ctx->devs[0];
        d.rc = 0;
        cudaSetDevice(d.id);
        int rc = fn(d, (size_t)0, n);
        if (rc) ctx->err = d.err;
        return rc;
    }
    std::vector<std::thread> th;
    for (size_t g = 0; g < nd; g++) {
        size_t lo = n * g / nd, hi = n * (g + 1) / nd;
        Dev *d = &ctx->devs[g];
        d->rc = 0;
        th.emplace_back([=]() {
            cudaSetDevice(d->id);
            d->rc = fn(*d, lo, hi);
        });
    }
    for (auto &t : th) t.join();
    for (auto &d : ctx->devs)
        if (d.rc) {
            ctx->err = d.err;
            return d.rc;
        }
    return 0;
}

int fail_arg(bn254_ctx *ctx, const char *what) {
    if (ctx) ctx->err = what;
    return BN254_ERR_ARG;
}

// ---- chunk pipeline of the host-buffer entry points ---------------------------------------------------
// A 2^24-tuple job is 64 chunks of 2^18 tuples per device.  Copies run on the copy stream d.cp, kernels on d.st, with two
// sets of staging buffers: while the kernels of chunk k read set k & 1, chunk k+1 is copied into the other set, and the
// results of chunk k-1 travel back.  Result bytes go straight to the caller's buffer when it is pinned (cudaHostAlloc /
// cudaHostRegister / bn254_host_register: the copy is then truly asynchronous); a pageable result buffer is served through
// a pinned staging area of the ctx, because a device-to-pageable copy would block the host for the whole chunk and with
// it the upload of the next one.  Pageable INPUT only costs the host-side staging copy of cudaMemcpyAsync, which overlaps
// the kernels already queued.
bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
struct Pipe {
    struct Pending { uint8_t *host; size_t off, bytes; };
    Dev &d;
    size_t k = 0;
    size_t pin_used = 0;
    std::vector<Pending> pend[2];
    cudaStream_t rs;      // the stream a chunk's results appear on (the launch stream unless a tail stream finishes the chunk)
    explicit Pipe(Dev &dd) : d(dd), rs(dd.st) {}
    int set() const { return (int)(k & 1); }
    // staging set k & 1 was last used by chunk k - 2: wait for it and hand its staged results to the caller
    int deliver(int s) {
        if (pend[s].empty() && k < 2) return 0;
        CK(d, cudaEventSynchronize(d.ev_out[s]));
        for (auto &q : pend[s]) memcpy(q.host, (const uint8_t *)d.pin[s] + q.off, q.bytes);
        pend[s].clear();
        return 0;
    }
    int begin(size_t staged_result_bytes) {
        int s = set();
        if (k >= 2)
            if (int rc = deliver(s)) return rc;
        if (d.pin_cap[s] < staged_result_bytes) {
            if (d.pin[s]) CK(d, cudaFreeHost(d.pin[s]));
            d.pin[s] = nullptr;
            d.pin_cap[s] = 0;
            size_t cap = staged_result_bytes + staged_result_bytes / 8 + 4096;
            CK(d, cudaHostAlloc(&d.pin[s], cap, cudaHostAllocDefault));
            d.pin_cap[s] = cap;
        }
        pin_used = 0;
        return 0;
    }
    int h2d(void *dst, const void *src, size_t bytes) {
        if (bytes) CK(d, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d.cp));
        return 0;
    }
    int inputs_ready() {
        int s = set();
        CK(d, cudaEventRecord(d.ev_in[s], d.cp));
        CK(d, cudaStreamWaitEvent(d.st, d.ev_in[s], 0));
        return 0;
    }
    // what has been copied since inputs_ready(): recorded, NOT yet waited for -- the consumer waits right before it needs it
    int more_inputs(cudaEvent_t *ev) {
        int s = set();
        CK(d, cudaEventRecord(d.ev_in2[s], d.cp));
        *ev = d.ev_in2[s];
        return 0;
    }
    int d2h(uint8_t *host, bool pinned, const uint8_t *dev, size_t bytes) {
        if (!bytes) return 0;
        if (pinned) {
            CK(d, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, rs));
            return 0;
        }
        int s = set();
        CK(d, cudaMemcpyAsync((uint8_t *)d.pin[s] + pin_used, dev, bytes, cudaMemcpyDeviceToHost, rs));
        pend[s].push_back({host, pin_used, bytes});
        pin_used += (bytes + 63) & ~(size_t)63;
        return 0;
    }
    int end() {
        CK(d, cudaEventRecord(d.ev_out[set()], rs));
        k++;
        return 0;
    }
    int finish() {
        // the last two chunks, oldest first
        for (int j = 0; j < 2; j++) {
            int s = (int)((k + j) & 1);
            if (k + j >= 2)
                if (int rc = deliver(s)) return rc;
        }
        CK(d, cudaStreamSynchronize(d.st));
        if (rs != d.st) CK(d, cudaStreamSynchronize(rs));
        return 0;
    }
};
inline size_t pad64(size_t b) { return (b + 63) & ~(size_t)63; }

// stage messages [lo, hi) of a host blob into staging set `set` on the copy stream: the byte range and the offsets
int stage_msgs(Dev &d, Pipe *pipe, const uint8_t *msgs, const uint64_t *offsets, size_t lo, size_t hi, const uint8_t **d_msgs,
               const uint64_t **d_off, uint64_t *base) {
    uint64_t b0 = offsets[lo], b1 = offsets[hi];
    uint8_t *dm; uint64_t *dof;
    int set = pipe ? pipe->set() : 0;
    cudaStream_t st = pipe ? d.cp : d.st;
    if (int rc = ensure_t(d, stage_buf(S_MSGS, set), (size_t)(b1 - b0) + 16, &dm)) return rc;
    if (int rc = ensure_t(d, stage_buf(S_OFF, set), hi - lo + 1, &dof)) return rc;
    if (b1 > b0) CK(d, cudaMemcpyAsync(dm, msgs + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
    CK(d, cudaMemcpyAsync(dof, offsets + lo, (hi - lo + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    *d_msgs = dm;
    *d_off = dof;
    *base = b0;
    return 0;
}

int check_offsets(bn254_ctx *ctx, const uint64_t *offsets, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail_arg(ctx, "offsets must be non-decreasing");
    return 0;
}

// host entry for the four hash/encode calls
int host_hash(bn254_ctx *ctx, int group, int count, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst,
              size_t dst_len, uint8_t *out, uint8_t *status) {
    if (!ctx) return BN254_ERR_ARG;
    if (n && (!offsets || !out)) return fail_arg(ctx, "null pointer");
    if (n == 0) return 0;
    if (int rc = check_offsets(ctx, offsets, n)) return rc;
    size_t rec = group == 1 ? 64 : 128;
    size_t chunk = ctx->chunk;
    const bool pin_out = host_ptr_is_pinned(out), pin_st = status && host_ptr_is_pinned(status);
    return run_sharded(ctx, n, [=](Dev &d, size_t lo, size_t hi) -> int {
        const uint8_t *ddst; uint32_t dlen;
        if (int rc = stage_dst(d, dst, dst_len, true, &ddst, &dlen)) return rc;
        Pipe pipe(d);
        for (size_t c = lo; c < hi; c += chunk) {
            size_t e = std::min(hi, c + chunk), m = e - c;
            if (int rc = pipe.begin((pin_out ? 0 : pad64(m * rec)) + (pin_st ? 0 : pad64(m)))) return rc;
            const uint8_t *dm; const uint64_t *dof; uint64_t base;
            if (int rc = stage_msgs(d, &pipe, msgs, offsets, c, e, &dm, &dof, &base)) return rc;
            if (int rc = pipe.inputs_ready()) return rc;
            uint8_t *dout, *dst_status;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), m * rec, &dout)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_STATUS, pipe.set()), m, &dst_status)) return rc;
            if (int rc = dev_hash_wire(d, group, dm, dof, base, m, ddst, dlen, count, dout, dst_status)) return rc;
            if (int rc = pipe.d2h(out + c * rec, pin_out, dout, m * rec)) return rc;
            if (status)
                if (int rc = pipe.d2h(status + c, pin_st, dst_status, m)) return rc;
            if (int rc = pipe.end()) return rc;
        }
        return pipe.finish();
    });
}

// generic "wire records in -> wire records out" host wrapper for point-wise kernels
template <class Fn>
int host_map(bn254_ctx *ctx, size_t n, const uint8_t *in_a, size_t rec_a, const uint8_t *in_b, size_t rec_b, uint8_t *out, size_t rec_out,
             uint8_t *status, Fn fn) {
    if (!ctx) return BN254_ERR_ARG;
    if (n == 0) return 0;
    if (!in_a || !out) return fail_arg(ctx, "null pointer");
    size_t chunk = ctx->chunk;
    const bool pin_out = host_ptr_is_pinned(out), pin_st = status && host_ptr_is_pinned(status);
    return run_sharded(ctx, n, [=](Dev &d, size_t lo, size_t hi) -> int {
        Pipe pipe(d);
        for (size_t c = lo; c < hi; c += chunk) {
            size_t e = std::min(hi, c + chunk), m = e - c;
            if (int rc = pipe.begin((pin_out ? 0 : pad64(m * rec_out)) + (pin_st ? 0 : pad64(m)))) return rc;
            uint8_t *da, *db = nullptr, *dout, *dstat;
            if (int rc = ensure_t(d, stage_buf(S_IN_A, pipe.set()), m * rec_a, &da)) return rc;
            if (int rc = pipe.h2d(da, in_a + c * rec_a, m * rec_a)) return rc;
            if (in_b) {
                if (int rc = ensure_t(d, stage_buf(S_IN_B, pipe.set()), m * rec_b, &db)) return rc;
                if (int rc = pipe.h2d(db, in_b + c * rec_b, m * rec_b)) return rc;
            }
            if (int rc = pipe.inputs_ready()) return rc;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), m * rec_out, &dout)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_STATUS, pipe.set()), m, &dstat)) return rc;
            if (int rc = fn(d, da, db, m, dout, dstat)) return rc;
            if (int rc = pipe.d2h(out + c * rec_out, pin_out, dout, m * rec_out)) return rc;
            if (status)
                if (int rc = pipe.d2h(status + c, pin_st, dstat, m)) return rc;
            if (int rc = pipe.end()) return rc;
        }
        return pipe.finish();
    });
}

}  // namespace

// =====================================================================================================
extern "C" {

int bn254_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bn254_ctx_create(bn254_ctx **out, const int *device_ids, int n_dev) {
    if (!out || n_dev < 1) return BN254_ERR_ARG;
    *out = nullptr;
    int have = bn254_device_count();
    if (have < 1) return BN254_ERR_NO_DEVICE;
    bn254_ctx *ctx = new bn254_ctx();
    for (int i = 0; i < n_dev; i++) {
        Dev d;
        d.id = device_ids ? device_ids[i] : i;
        if (d.id < 0 || d.id >= have) {
            delete ctx;
            return BN254_ERR_ARG;
        }
        ctx->devs.push_back(d);
    }
    for (auto &d : ctx->devs) {
        cudaError_t e = cudaSetDevice(d.id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking);
        d.cur = d.st;
        if (e == cudaSuccess) {
            cudaDeviceProp prop;
            e = cudaGetDeviceProperties(&prop, d.id);
            d.sms = (size_t)prop.multiProcessorCount;
            d.lane_wave_threads = d.sms * BN_MINB_L2 * BN_TPB_L2;
        }
        for (int k = 0; k < Dev::MAXSUB && e == cudaSuccess; k++) {
            e = cudaStreamCreateWithFlags(&d.sub[k], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_join[k], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) {
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            e = cudaStreamCreateWithPriority(&d.hi, cudaStreamNonBlocking, hi_pri);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_hi_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_hi_join, cudaEventDisableTiming);
        if (e == cudaSuccess) {
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            e = cudaStreamCreateWithPriority(&d.tail, cudaStreamNonBlocking, hi_pri);
        }
        for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&d.ev_tail[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.cp, cudaStreamNonBlocking);
        for (int k = 0; k < 2 && e == cudaSuccess; k++) {
            e = cudaEventCreateWithFlags(&d.ev_in[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_in2[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_out[k], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            int rc = BN254_ERR_CUDA_BASE - (int)e;
            bn254_ctx_destroy(ctx);
            return rc;
        }
    }
    for (auto &d : ctx->devs) {   // the slot-machine kernels live on shared memory: ask for the largest carve-out
        cudaSetDevice(d.id);
        cudaFuncSetAttribute(k_vm_miller2_fixed, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_vm_miller2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_vm_miller1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_vm_final_exp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (const char *s = getenv("BN254_CARVEOUT")) {  // experiment knob: preferred shared-memory carveout (percent) of the pairing kernels
        int v = atoi(s);
        for (auto &d : ctx->devs) {
            cudaSetDevice(d.id);
            cudaFuncSetAttribute(k_miller2_fixed, cudaFuncAttributePreferredSharedMemoryCarveout, v);
            cudaFuncSetAttribute(k_miller2_fixed_l2, cudaFuncAttributePreferredSharedMemoryCarveout, v);
            cudaFuncSetAttribute(k_final_exp, cudaFuncAttributePreferredSharedMemoryCarveout, v);
            cudaFuncSetAttribute(k_final_exp_l2, cudaFuncAttributePreferredSharedMemoryCarveout, v);
            cudaFuncSetAttribute(k_hash_g1, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        }
    }
    if (const char *s = getenv("BN254_STACK")) {  // bring-up knob: per-thread stack limit in bytes
        long v = atol(s);
        for (auto &d : ctx->devs) {
            cudaSetDevice(d.id);
            if (v > 0) cudaDeviceSetLimit(cudaLimitStackSize, (size_t)v);
        }
    }
    if (const char *s = getenv("BN254_SPLIT")) {  // slices per verification batch (1 = no concurrency)
        long v = atol(s);
        if (v >= 1 && v <= Dev::MAXSUB)
            for (auto &d : ctx->devs) d.nsub = (int)v;
    }
    if (const char *s = getenv("BN254_CHUNK")) {
        long v = atol(s);
        if (v > 0) {
            ctx->chunk = (size_t)v;
            ctx->chunk_is_default = false;
        }
    }
    if (const char *s = getenv("BN254_LANES")) {  // shape of the pairing kernels: 0 per launch, 1 one thread, 2 lane pair, 3 / 4 slot machines
        long v = atol(s);
        if (v >= 0 && v <= 4)
            for (auto &d : ctx->devs) d.shape = (int)v;
    }
    if (const char *s = getenv("BN254_WIDE_MAX")) {
        long v = atol(s);
        if (v >= 0)
            for (auto &d : ctx->devs) d.wide_max = (size_t)v;
    }
    // Several devices in one process: one NCCL communicator per device for the aggregate path's all-gather.  A box without
    // libnccl still gets a working ctx for everything that needs no exchange; the aggregate calls then fail with -2000.
    if (ctx->devs.size() > 1 && bn_nccl::api().ok()) {
        std::vector<bn_nccl::comm_t> comms(ctx->devs.size());
        std::vector<int> ids;
        for (auto &d : ctx->devs) ids.push_back(d.id);
        int r = bn_nccl::api().CommInitAll(comms.data(), (int)ids.size(), ids.data());
        if (r != bn_nccl::kSuccess) {
            bn254_ctx_destroy(ctx);
            return BN254_ERR_NCCL_BASE - r;
        }
        for (size_t g = 0; g < ctx->devs.size(); g++) {
            ctx->devs[g].comm = comms[g];
            ctx->devs[g].nranks = (int)ctx->devs.size();
            ctx->devs[g].rank = (int)g;
        }
    }
    *out = ctx;
    return 0;
}

int bn254_ctx_set_option(bn254_ctx *ctx, const char *name, long value) {
    if (!ctx || !name) return BN254_ERR_ARG;
    std::string k(name);
    if (k == "lanes" && value >= 0 && value <= 4) {
        for (auto &d : ctx->devs) d.shape = (int)value;
    } else if (k == "wide_max" && value >= 0) {
        for (auto &d : ctx->devs) d.wide_max = (size_t)value;
    } else if (k == "chunk" && value > 0) {
        ctx->chunk = (size_t)value;
        ctx->chunk_is_default = false;
    } else if (k == "sk_msm_slices" && value >= 0 && value <= 64) {
        for (auto &d : ctx->devs) d.sk_msm_slices = (size_t)value;
    } else if (k == "sk_msm_min" && value >= 0) {
        for (auto &d : ctx->devs) d.sk_msm_min = (size_t)value;
    } else if (k == "split" && value >= 1 && value <= Dev::MAXSUB) {
        for (auto &d : ctx->devs) d.nsub = (int)value;
    } else {
        return fail_arg(ctx, "unknown option or value out of range (lanes 0..4, wide_max >= 0, chunk > 0, split 1..8, sk_msm_min >= 0)");
    }
    return 0;
}

// ---- NCCL communicator of a one-device ctx (one process per GPU: torchrun, MPI, ...) -------------------------------
int bn254_comm_unique_id(uint8_t *out) {
    if (!out) return BN254_ERR_ARG;
    if (!bn_nccl::api().ok()) return BN254_ERR_NCCL_BASE;
    bn_nccl::unique_id id;
    int r = bn_nccl::api().GetUniqueId(&id);
    if (r != bn_nccl::kSuccess) return BN254_ERR_NCCL_BASE - r;
    memcpy(out, id.internal, BN254_COMM_ID_BYTES);
    return 0;
}
int bn254_comm_init_rank(bn254_ctx *ctx, const uint8_t *id_bytes, int nranks, int rank) {
    if (!ctx || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return fail_arg(ctx, "bad communicator arguments");
    if (ctx->devs.size() != 1) return fail_arg(ctx, "bn254_comm_init_rank needs a one-device ctx (a multi-device ctx owns its communicators)");
    Dev &d = ctx->devs[0];
    if (d.comm) return fail_arg(ctx, "communicator already initialised");
    if (!bn_nccl::api().ok()) {
        ctx->err = bn_nccl::api().why ? bn_nccl::api().why : "NCCL unavailable";
        return BN254_ERR_NCCL_BASE;
    }
    cudaSetDevice(d.id);
    bn_nccl::unique_id id;
    memcpy(id.internal, id_bytes, BN254_COMM_ID_BYTES);
    int r = bn_nccl::api().CommInitRank(&d.comm, nranks, id, rank);
    if (r != bn_nccl::kSuccess) {
        d.comm = nullptr;
        ctx->err = std::string("ncclCommInitRank: ") + bn_nccl::api().GetErrorString(r);
        return BN254_ERR_NCCL_BASE - r;
    }
    d.nranks = nranks;
    d.rank = rank;
    return 0;
}
int bn254_comm_nranks(const bn254_ctx *ctx) { return ctx ? ctx->devs[0].nranks : 0; }

// pin / unpin caller memory so that the chunk pipeline's copies are asynchronous (cudaHostRegister)
int bn254_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return BN254_ERR_ARG;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    return e == cudaSuccess ? 0 : BN254_ERR_CUDA_BASE - (int)e;
}
int bn254_host_unregister(void *p) {
    if (!p) return BN254_ERR_ARG;
    cudaError_t e = cudaHostUnregister(p);
    return e == cudaSuccess ? 0 : BN254_ERR_CUDA_BASE - (int)e;
}

void bn254_ctx_destroy(bn254_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.id);
        if (d.st) cudaStreamSynchronize(d.st);
        if (d.hi) cudaStreamSynchronize(d.hi);   // joined into d.st by every call; an aborted call may have left work on it
        for (auto &b : d.bufs)
            if (b.p) cudaFree(b.p);
        for (int k = 0; k < Dev::MAXSUB; k++) {
            if (d.sub[k]) cudaStreamDestroy(d.sub[k]);
            if (d.ev_join[k]) cudaEventDestroy(d.ev_join[k]);
        }
        if (d.ev_fork) cudaEventDestroy(d.ev_fork);
        if (d.hi) cudaStreamDestroy(d.hi);
        if (d.ev_hi_fork) cudaEventDestroy(d.ev_hi_fork);
        if (d.ev_hi_join) cudaEventDestroy(d.ev_hi_join);
        if (d.cp) {
            cudaStreamSynchronize(d.cp);
            cudaStreamDestroy(d.cp);
        }
        if (d.tail) {
            cudaStreamSynchronize(d.tail);
            cudaStreamDestroy(d.tail);
        }
        for (int k = 0; k < 2; k++)
            if (d.ev_tail[k]) cudaEventDestroy(d.ev_tail[k]);
        for (int k = 0; k < 2; k++) {
            if (d.ev_in[k]) cudaEventDestroy(d.ev_in[k]);
            if (d.ev_in2[k]) cudaEventDestroy(d.ev_in2[k]);
            if (d.ev_out[k]) cudaEventDestroy(d.ev_out[k]);
            if (d.pin[k]) cudaFreeHost(d.pin[k]);
        }
        for (auto &r : d.recs) {
            if (r.e0) cudaEventDestroy(r.e0);
            if (r.e1) cudaEventDestroy(r.e1);
        }
        if (d.comm && bn_nccl::api().ok()) bn_nccl::api().CommDestroy(d.comm);
        if (d.st) cudaStreamDestroy(d.st);
    }
    delete ctx;
}

const char *bn254_last_error(const bn254_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

uint64_t bn254_launch_count(const bn254_ctx *ctx) {
    uint64_t t = 0;
    if (ctx)
        for (auto &d : ctx->devs) t += d.launches;
    return t;
}
int bn254_profile_enable(bn254_ctx *ctx, int on) {
    if (!ctx) return BN254_ERR_ARG;
    for (auto &d : ctx->devs) d.prof = on != 0;
    return 0;
}
// Drains the recorded launches of every device (device order) into `buf` as lines "kernel items ms\n".
int bn254_profile_drain(bn254_ctx *ctx, char *buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return BN254_ERR_ARG;
    std::string out;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.id);
        cudaStreamSynchronize(d.st);
        for (auto &r : d.recs) {
            float ms = 0;
            if (r.e0 && r.e1 && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
                char line[160];
                snprintf(line, sizeof line, "%s %zu %.6f\n", r.name, r.items, ms);
                out += line;
            }
            if (r.e0) cudaEventDestroy(r.e0);
            if (r.e1) cudaEventDestroy(r.e1);
        }
        d.recs.clear();
    }
    if (out.size() + 1 > cap) return BN254_ERR_ARG;
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}
void *bn254_stream(bn254_ctx *ctx) { return ctx ? (void *)ctx->devs[0].st : nullptr; }
int bn254_sync(bn254_ctx *ctx) {
    if (!ctx) return BN254_ERR_ARG;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.id);
        cudaError_t e = cudaStreamSynchronize(d.st);
        if (e != cudaSuccess) {
            ctx->err = cudaGetErrorString(e);
            return BN254_ERR_CUDA_BASE - (int)e;
        }
    }
    return 0;
}

int bn254_hash_to_g1(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, uint8_t *out,
                     uint8_t *status) {
    return host_hash(ctx, 1, 2, msgs, offsets, n, dst, dst_len, out, status);
}
int bn254_encode_to_g1(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                       uint8_t *out, uint8_t *status) {
    return host_hash(ctx, 1, 1, msgs, offsets, n, dst, dst_len, out, status);
}
int bn254_hash_to_g2(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, uint8_t *out,
                     uint8_t *status) {
    return host_hash(ctx, 2, 2, msgs, offsets, n, dst, dst_len, out, status);
}
int bn254_encode_to_g2(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                       uint8_t *out, uint8_t *status) {
    return host_hash(ctx, 2, 1, msgs, offsets, n, dst, dst_len, out, status);
}

int bn254_pairing_batch(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *gt, uint8_t *status) {
    if (n && !g2) return fail_arg(ctx, "null pointer");
    return host_map(ctx, n, g1, 64, g2, 128, gt, 384, status,
                    [](Dev &d, const uint8_t *a, const uint8_t *b, size_t m, uint8_t *o, uint8_t *s) { return dev_pairing(d, a, b, m, o, s, 0); });
}
int bn254_miller_loop_batch(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *ml, uint8_t *status) {
    if (n && !g2) return fail_arg(ctx, "null pointer");
    return host_map(ctx, n, g1, 64, g2, 128, ml, 384, status,
                    [](Dev &d, const uint8_t *a, const uint8_t *b, size_t m, uint8_t *o, uint8_t *s) { return dev_pairing(d, a, b, m, o, s, 1); });
}
int bn254_final_exponentiation(bn254_ctx *ctx, const uint8_t *ml, size_t n, uint8_t *gt, uint8_t *status) {
    return host_map(ctx, n, ml, 384, nullptr, 0, gt, 384, status,
                    [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *s) { return dev_final_exp(d, a, m, o, s); });
}

int bn254_fp12_product(bn254_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out) {
    if (!ctx || !out || (n && !in)) return fail_arg(ctx, "null pointer");
    Dev &d = ctx->devs[0];
    d.rc = 0;
    cudaSetDevice(d.id);
    auto body = [&]() -> int {
        uint8_t *din, *dout, *flf; fp12 *F; int *dbad;
        if (int rc = ensure_t(d, S_IN_A, n * 384 + 16, &din)) return rc;
        if (int rc = ensure_t(d, S_OUT, 384, &dout)) return rc;
        if (int rc = ensure_t(d, S_F, n + 1, &F)) return rc;
        if (int rc = ensure_t(d, S_FLF, n + 1, &flf)) return rc;
        if (int rc = ensure_t(d, S_BAD, 4, &dbad)) return rc;
        CK(d, cudaMemsetAsync(dbad, 0, sizeof(int), d.st));
        if (n) CK(d, cudaMemcpyAsync(din, in, n * 384, cudaMemcpyHostToDevice, d.st));
        LAUNCH(d, k_fp12_decode, n, BN_TPB_LIGHT, F, flf, din, n);
        LAUNCH(d, k_or_flags, n, BN_TPB_LIGHT, dbad, flf, n);
        fp12 *res;
        if (int rc = dev_fp12_product(d, F, n, &res)) return rc;
        LAUNCH(d, k_fp12_encode, 1, BN_TPB_LIGHT, dout, (uint8_t *)nullptr, res, (const uint8_t *)nullptr, (size_t)1);
        int hbad = 0;
        CK(d, cudaMemcpyAsync(out, dout, 384, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaMemcpyAsync(&hbad, dbad, sizeof(int), cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        return hbad ? BN254_ERR_INVALID_GT_BYTES : 0;
    };
    int rc = body();
    if (rc && rc != BN254_ERR_INVALID_GT_BYTES) ctx->err = d.err;
    return rc;
}

int bn254_multi_miller_loop(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *ml) {
    if (!ctx || !ml || (n && (!g1 || !g2))) return fail_arg(ctx, "null pointer");
    Dev &d = ctx->devs[0];
    d.rc = 0;
    cudaSetDevice(d.id);
    int which = 0;
    auto body = [&]() -> int {
        uint8_t *da, *db, *dout, *flp, *flq; g1m *P; g2m *Q; fp12 *F; int *dbad;
        if (int rc = ensure_t(d, S_IN_A, n * 64 + 16, &da)) return rc;
        if (int rc = ensure_t(d, S_IN_B, n * 128 + 16, &db)) return rc;
        if (int rc = ensure_t(d, S_OUT, 384, &dout)) return rc;
        if (int rc = ensure_t(d, S_P0, n + 1, &P)) return rc;
        if (int rc = ensure_t(d, S_Q0, n + 1, &Q)) return rc;
        if (int rc = ensure_t(d, S_FLP0, n + 1, &flp)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, n + 1, &flq)) return rc;
        if (int rc = ensure_t(d, S_F, n + 1, &F)) return rc;
        if (int rc = ensure_t(d, S_BAD, 8, &dbad)) return rc;
        CK(d, cudaMemsetAsync(dbad, 0, 2 * sizeof(int), d.st));
        if (n) {
            CK(d, cudaMemcpyAsync(da, g1, n * 64, cudaMemcpyHostToDevice, d.st));
            CK(d, cudaMemcpyAsync(db, g2, n * 128, cudaMemcpyHostToDevice, d.st));
        }
        LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P, flp, da, n, 0);
        LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, Q, flq, db, n, 0);
        LAUNCH(d, k_or_flags, n, BN_TPB_LIGHT, dbad, flp, n);
        LAUNCH(d, k_or_flags, n, BN_TPB_LIGHT, dbad + 1, flq, n);
        if (int rc = launch_miller1(d, n, F, (uint8_t *)nullptr, P, flp, Q, flq, (int *)nullptr)) return rc;
        fp12 *res;
        if (int rc = dev_fp12_product(d, F, n, &res)) return rc;
        LAUNCH(d, k_fp12_encode, 1, BN_TPB_LIGHT, dout, (uint8_t *)nullptr, res, (const uint8_t *)nullptr, (size_t)1);
        int hbad[2] = {0, 0};
        CK(d, cudaMemcpyAsync(ml, dout, 384, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaMemcpyAsync(hbad, dbad, 2 * sizeof(int), cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        which = hbad[0] ? BN254_ERR_INVALID_G1_BYTES : (hbad[1] ? BN254_ERR_INVALID_G2_BYTES : 0);
        return which;
    };
    int rc = body();
    if (rc && !which) ctx->err = d.err;
    return rc;
}

// ---- verify ------------------------------------------------------------------------------------
int bn254_verify_batch(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, const uint8_t *sigs,
                       size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    if (!ctx) return BN254_ERR_ARG;
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (n == 0) return 0;
    if (!pks || !offsets || !sigs || !ok) return fail_arg(ctx, "null pointer");
    if (int rc = check_offsets(ctx, offsets, n)) return rc;
    size_t rpk = variant == BN254_MIN_SIG ? 128 : 64, rsig = variant == BN254_MIN_SIG ? 64 : 128;
    const bool pin_ok = host_ptr_is_pinned(ok);
    return run_sharded(ctx, n, [=](Dev &d, size_t lo, size_t hi) -> int {
        const uint8_t *ddst; uint32_t dlen;
        if (int rc = stage_dst(d, dst, dst_len, true, &ddst, &dlen)) return rc;
        const size_t chunk = ctx->verify_chunk(d);
        Pipe pipe(d);
        for (size_t c = lo; c < hi; c += chunk) {
            size_t e = std::min(hi, c + chunk), m = e - c;
            if (int rc = pipe.begin(pin_ok ? 0 : pad64(m))) return rc;
            const uint8_t *dm; const uint64_t *dof; uint64_t base;
            if (int rc = stage_msgs(d, &pipe, msgs, offsets, c, e, &dm, &dof, &base)) return rc;
            uint8_t *dpk, *dsig, *dok;
            if (int rc = ensure_t(d, stage_buf(S_IN_A, pipe.set()), m * rpk, &dpk)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_B, pipe.set()), m * rsig, &dsig)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), m, &dok)) return rc;
            if (int rc = pipe.inputs_ready()) return rc;       // the messages: all the hash kernel needs
            if (int rc = pipe.h2d(dpk, pks + c * rpk, m * rpk)) return rc;
            if (int rc = pipe.h2d(dsig, sigs + c * rsig, m * rsig)) return rc;
            cudaEvent_t rest;
            if (int rc = pipe.more_inputs(&rest)) return rc;
            if (int rc = dev_verify(d, variant, dpk, dm, dof, base, dsig, m, ddst, dlen, flags, dok, rest)) return rc;
            if (int rc = pipe.d2h(ok + c, pin_ok, dok, m)) return rc;
            if (int rc = pipe.end()) return rc;
        }
        return pipe.finish();
    });
}

int bn254_verify_batch_randomized(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets,
                                  const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, const uint8_t *seed,
                                  int rand_bits, size_t group, uint8_t *ok_groups) {
    if (!ctx) return BN254_ERR_ARG;
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (rand_bits != 64 && rand_bits != 128) return fail_arg(ctx, "rand_bits must be 64 or 128");
    if (group == 0 || group > BN254_RBV_MAX_GROUP) return fail_arg(ctx, "group must be in 1..BN254_RBV_MAX_GROUP");
    if (n == 0) return 0;
    if (!pks || !offsets || !sigs || !seed || !ok_groups) return fail_arg(ctx, "null pointer");
    if (int rc = check_offsets(ctx, offsets, n)) return rc;
    size_t rpk = variant == BN254_MIN_SIG ? 128 : 64, rsig = variant == BN254_MIN_SIG ? 64 : 128;
    size_t nseg = (n + group - 1) / group;
    size_t chunk_groups = std::max<size_t>(1, ctx->chunk / group);   // a group never straddles two chunks
    const bool pin_ok = host_ptr_is_pinned(ok_groups);
    // groups (not tuples) are what the devices share out
    return run_sharded(ctx, nseg, [=](Dev &d, size_t glo, size_t ghi) -> int {
        const uint8_t *ddst; uint32_t dlen;
        if (int rc = stage_dst(d, dst, dst_len, true, &ddst, &dlen)) return rc;
        uint8_t *dseed;
        if (int rc = ensure_t(d, S_RB_SEED, 32, &dseed)) return rc;
        CK(d, cudaMemcpyAsync(dseed, seed, 32, cudaMemcpyHostToDevice, d.st));
        Pipe pipe(d);
        const bool overlap_tail = !d.prof && d.shape != 3;
        if (overlap_tail) pipe.rs = d.tail;
        for (size_t g = glo; g < ghi; g += chunk_groups) {
            size_t ge = std::min(ghi, g + chunk_groups);
            size_t c = g * group, e = std::min(n, ge * group), m = e - c;
            if (int rc = pipe.begin(pin_ok ? 0 : pad64(ge - g))) return rc;
            const uint8_t *dm; const uint64_t *dof; uint64_t base;
            if (int rc = stage_msgs(d, &pipe, msgs, offsets, c, e, &dm, &dof, &base)) return rc;
            uint8_t *dpk, *dsig, *dok;
            if (int rc = ensure_t(d, stage_buf(S_IN_A, pipe.set()), m * rpk, &dpk)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_B, pipe.set()), m * rsig, &dsig)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), ge - g, &dok)) return rc;
            if (int rc = pipe.h2d(dpk, pks + c * rpk, m * rpk)) return rc;
            if (int rc = pipe.h2d(dsig, sigs + c * rsig, m * rsig)) return rc;
            if (int rc = pipe.inputs_ready()) return rc;
            if (int rc = dev_verify_rand(d, variant, dpk, dm, dof, base, dsig, m, group, dseed, (uint64_t)c, rand_bits, ddst, dlen, flags, dok,
                                         overlap_tail ? pipe.set() : -1, ge < ghi /* another chunk follows and hides this tail */))
                return rc;
            if (int rc = pipe.d2h(ok_groups + g, pin_ok, dok, ge - g)) return rc;
            if (int rc = pipe.end()) return rc;
        }
        return pipe.finish();
    });
}

// Same-key randomised verification: key k owns the tuples [key_offsets[k], key_offsets[k+1]) of msgs / sigs and gets one accept
// byte.  Keys (with their tuples) are what the devices share out and what the chunk pipeline streams; a key is never split, so
// the scratch grows to the largest key (~0.5 KB of device memory per tuple).
int bn254_verify_same_key_randomized(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint64_t *key_offsets, size_t n_keys,
                                     const uint8_t *msgs, const uint64_t *offsets, const uint8_t *sigs, const uint8_t *dst, size_t dst_len,
                                     uint32_t flags, const uint8_t *seed, int rand_bits, uint8_t *ok_keys) {
    if (!ctx) return BN254_ERR_ARG;
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (rand_bits != 64 && rand_bits != 128) return fail_arg(ctx, "rand_bits must be 64 or 128");
    if (n_keys == 0) return 0;
    if (!pks || !key_offsets || !seed || !ok_keys) return fail_arg(ctx, "null pointer");
    if (int rc = check_offsets(ctx, key_offsets, n_keys)) return rc;
    const size_t n_tuples = (size_t)key_offsets[n_keys];
    if (n_tuples && (!offsets || !sigs)) return fail_arg(ctx, "null pointer");
    if (n_tuples)
        if (int rc = check_offsets(ctx, offsets, n_tuples)) return rc;
    const size_t rpk = variant == BN254_MIN_SIG ? 128 : 64, rsig = variant == BN254_MIN_SIG ? 64 : 128;
    const bool pin_ok = host_ptr_is_pinned(ok_keys);
    return run_sharded(ctx, n_keys, [=](Dev &d, size_t lo, size_t hi) -> int {
        const uint8_t *ddst; uint32_t dlen;
        if (int rc = stage_dst(d, dst, dst_len, true, &ddst, &dlen)) return rc;
        uint8_t *dseed;
        if (int rc = ensure_t(d, S_RB_SEED, 32, &dseed)) return rc;
        CK(d, cudaMemcpyAsync(dseed, seed, 32, cudaMemcpyHostToDevice, d.st));
        Pipe pipe(d);
        const bool overlap_tail = !d.prof && d.shape != 3;
        if (overlap_tail) pipe.rs = d.tail;
        SkPlan plan;
        for (size_t c = lo; c < hi;) {
            size_t e = c + 1;
            while (e < hi && e - c < ctx->chunk && key_offsets[e + 1] - key_offsets[c] <= ctx->chunk) e++;
            const size_t nk = e - c, t0 = (size_t)key_offsets[c], m = (size_t)key_offsets[e] - t0;
            plan = SkPlan();
            sk_plan(plan, key_offsets + c, nk, !sk_use_msm(d, m, nk));
            if (int rc = pipe.begin(pin_ok ? 0 : pad64(nk))) return rc;
            const uint8_t *dm = nullptr; const uint64_t *dof = nullptr; uint64_t base = 0;
            if (m)
                if (int rc = stage_msgs(d, &pipe, msgs, offsets, t0, t0 + m, &dm, &dof, &base)) return rc;
            uint8_t *dpk, *dsig, *dok; uint64_t *dplan;
            if (int rc = ensure_t(d, stage_buf(S_IN_A, pipe.set()), nk * rpk, &dpk)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_B, pipe.set()), m * rsig + 16, &dsig)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_C, pipe.set()), plan.blob.size(), &dplan)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), nk, &dok)) return rc;
            if (int rc = pipe.h2d(dpk, pks + c * rpk, nk * rpk)) return rc;
            if (int rc = pipe.h2d(dsig, sigs + t0 * rsig, m * rsig)) return rc;
            // pageable source: the runtime stages it before the call returns, so the plan may be rebuilt for the next chunk
            if (int rc = pipe.h2d(dplan, plan.blob.data(), plan.blob.size() * sizeof(uint64_t))) return rc;
            if (int rc = pipe.inputs_ready()) return rc;
            if (int rc = dev_verify_same_key(d, variant, dpk, nk, dplan, plan, dm, dof, base, dsig, m, dseed, (uint64_t)t0, rand_bits, ddst, dlen,
                                             flags, dok, overlap_tail ? pipe.set() : -1, false /* the tail keeps the latency shape: see dev_verify_same_key */))
                return rc;
            if (int rc = pipe.d2h(ok_keys + c, pin_ok, dok, nk)) return rc;
            if (int rc = pipe.end()) return rc;
            c = e;
        }
        return pipe.finish();
    });
}

// FastAggregateVerify of the IETF BLS draft for G groups: group g = the public keys [key_offsets[g], key_offsets[g+1]) of `pks`,
// message g, aggregate signature g.  Groups (with their keys) are what the devices share out and what the chunk pipeline streams.
int bn254_fast_aggregate_verify_batch(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint64_t *key_offsets, const uint8_t *msgs,
                                      const uint64_t *offsets, const uint8_t *agg_sigs, size_t n_groups, const uint8_t *dst, size_t dst_len,
                                      uint32_t flags, uint8_t *ok) {
    if (!ctx) return BN254_ERR_ARG;
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (n_groups == 0) return 0;
    if (!key_offsets || !offsets || !agg_sigs || !ok || (key_offsets[n_groups] && !pks)) return fail_arg(ctx, "null pointer");
    if (int rc = check_offsets(ctx, offsets, n_groups)) return rc;
    if (int rc = check_offsets(ctx, key_offsets, n_groups)) return rc;
    const size_t rpk = variant == BN254_MIN_SIG ? 128 : 64, rsig = variant == BN254_MIN_SIG ? 64 : 128;
    const bool pin_ok = host_ptr_is_pinned(ok);
    const size_t max_keys = ctx->chunk;     // keys staged per round; a single bigger group still goes through alone
    return run_sharded(ctx, n_groups, [=](Dev &d, size_t lo, size_t hi) -> int {
        const uint8_t *ddst; uint32_t dlen;
        if (int rc = stage_dst(d, dst, dst_len, true, &ddst, &dlen)) return rc;
        Pipe pipe(d);
        for (size_t c = lo; c < hi;) {
            size_t e = c + 1;
            while (e < hi && e - c < ctx->verify_chunk(d) && key_offsets[e + 1] - key_offsets[c] <= max_keys) e++;
            const size_t m = e - c, nk = (size_t)(key_offsets[e] - key_offsets[c]);
            if (int rc = pipe.begin(pin_ok ? 0 : pad64(m))) return rc;
            const uint8_t *dm; const uint64_t *dof; uint64_t base;
            if (int rc = stage_msgs(d, &pipe, msgs, offsets, c, e, &dm, &dof, &base)) return rc;
            uint8_t *dpk, *dsig, *dok; uint64_t *dko;
            if (int rc = ensure_t(d, stage_buf(S_IN_A, pipe.set()), nk * rpk + 16, &dpk)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_B, pipe.set()), m * rsig, &dsig)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_C, pipe.set()), m + 1, &dko)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), m, &dok)) return rc;
            if (int rc = pipe.h2d(dpk, pks + (size_t)key_offsets[c] * rpk, nk * rpk)) return rc;
            if (int rc = pipe.h2d(dsig, agg_sigs + c * rsig, m * rsig)) return rc;
            if (int rc = pipe.h2d(dko, key_offsets + c, (m + 1) * sizeof(uint64_t))) return rc;
            if (int rc = pipe.inputs_ready()) return rc;
            if (int rc = dev_fast_aggregate(d, variant, dpk, nk, dko, key_offsets[c], dm, dof, base, dsig, m, ddst, dlen, flags, dok)) return rc;
            if (int rc = pipe.d2h(ok + c, pin_ok, dok, m)) return rc;
            if (int rc = pipe.end()) return rc;
            c = e;
        }
        return pipe.finish();
    });
}

// ---- prepared public keys (G2Prepared, pairings.rs:609-660): prepare once, verify many --------------------------------------
int bn254_g2_prepare(bn254_ctx *ctx, const uint8_t *pks, size_t n, uint32_t flags, bn254_g2_prepared **out) {
    if (!ctx || !out || (n && !pks)) return fail_arg(ctx, "null pointer");
    *out = nullptr;
    bn254_g2_prepared *pr = new bn254_g2_prepared();
    pr->ctx = ctx;
    pr->n = n;
    const int check = ((flags & (BN254_FLAG_CHECK_ON_CURVE | BN254_FLAG_CHECK_SUBGROUP)) ? DEC_ON_CURVE : 0) | DEC_NO_IDENTITY;
    const bool subgroup = (flags & BN254_FLAG_CHECK_SUBGROUP) != 0;
    int rc = 0;
    for (auto &d : ctx->devs) {     // every device holds the whole table: tuples are sharded, keys are not
        cudaSetDevice(d.id);
        d.rc = 0;
        rc = [&]() -> int {
            void *tab = nullptr, *fl = nullptr;
            CK(d, cudaMalloc(&tab, std::max<size_t>(1, n) * BN254_NUM_LINES * sizeof(line_t<fp2>)));
            pr->tab.push_back(tab);
            CK(d, cudaMalloc(&fl, std::max<size_t>(16, n)));
            pr->fl.push_back(fl);
            uint8_t *din; g2m *Q;
            void *pts = nullptr;
            CK(d, cudaMalloc(&pts, std::max<size_t>(1, n) * sizeof(g2m)));   // kept: small batches run the ordinary (wide) loop on them
            pr->pts.push_back(pts);
            Q = (g2m *)pts;
            if (int r = ensure_t(d, S_IN_A, n * 128 + 16, &din)) return r;
            if (n) CK(d, cudaMemcpyAsync(din, pks, n * 128, cudaMemcpyHostToDevice, d.st));
            LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, Q, (uint8_t *)fl, din, n, check);
            if (subgroup) LAUNCH(d, k_g2_require_subgroup, n, BN_TPB_HEAVY, (uint8_t *)fl, Q, n);
            LAUNCH(d, k_g2_prepare, n, BN_TPB_HEAVY, (line_t<fp2> *)tab, Q, (const uint8_t *)fl, n);
            CK(d, cudaStreamSynchronize(d.st));
            return 0;
        }();
        if (rc) {
            ctx->err = d.err;
            break;
        }
    }
    if (rc) {
        bn254_g2_prepared_free(pr);
        return rc;
    }
    *out = pr;
    return 0;
}
void bn254_g2_prepared_free(bn254_g2_prepared *pr) {
    if (!pr) return;
    for (size_t g = 0; g < pr->tab.size(); g++) {
        cudaSetDevice(pr->ctx->devs[g].id);
        if (pr->tab[g]) cudaFree(pr->tab[g]);
        if (g < pr->fl.size() && pr->fl[g]) cudaFree(pr->fl[g]);
        if (g < pr->pts.size() && pr->pts[g]) cudaFree(pr->pts[g]);
    }
    delete pr;
}
size_t bn254_g2_prepared_count(const bn254_g2_prepared *pr) { return pr ? pr->n : 0; }
int bn254_g2_prepared_export(bn254_ctx *ctx, const bn254_g2_prepared *pr, size_t first, size_t count, uint8_t *out, uint8_t *status) {
    if (!ctx || !pr || pr->ctx != ctx || !out || first + count > pr->n) return fail_arg(ctx, "bad prepared-key range");
    if (!count) return 0;
    Dev &d = ctx->devs[0];
    d.rc = 0;
    cudaSetDevice(d.id);
    auto body = [&]() -> int {
        const size_t nfp2 = count * BN254_NUM_LINES * 3;
        uint8_t *dout;
        if (int rc = ensure_t(d, S_OUT, nfp2 * 64, &dout)) return rc;
        LAUNCH(d, k_prepared_encode, nfp2, BN_TPB_LIGHT, dout, (const line_t<fp2> *)pr->tab[0] + first * BN254_NUM_LINES, nfp2);
        CK(d, cudaMemcpyAsync(out, dout, nfp2 * 64, cudaMemcpyDeviceToHost, d.st));
        std::vector<uint8_t> fl(count);
        CK(d, cudaMemcpyAsync(fl.data(), (const uint8_t *)pr->fl[0] + first, count, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        for (size_t i = 0; i < count; i++) {
            const bool good = !(fl[i] & (FL_BAD | FL_INF));
            if (status) status[i] = good ? 1 : 0;
            if (!good) memset(out + i * BN254_NUM_LINES * 192, 0, (size_t)BN254_NUM_LINES * 192);   // the table of an unusable key is undefined
        }
        return 0;
    };
    int rc = body();
    if (rc) ctx->err = d.err;
    return rc;
}
int bn254_verify_batch_prepared(bn254_ctx *ctx, const bn254_g2_prepared *pr, const uint32_t *key_index, const uint8_t *msgs, const uint64_t *offsets,
                                const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    if (!ctx) return BN254_ERR_ARG;
    if (!pr || pr->ctx != ctx) return fail_arg(ctx, "prepared keys belong to another ctx");
    if (n == 0) return 0;
    if (!key_index || !offsets || !sigs || !ok) return fail_arg(ctx, "null pointer");
    if (int rc = check_offsets(ctx, offsets, n)) return rc;
    for (size_t i = 0; i < n; i++)
        if (key_index[i] >= pr->n) return fail_arg(ctx, "key index out of range");
    const bool pin_ok = host_ptr_is_pinned(ok);
    return run_sharded(ctx, n, [=](Dev &d, size_t lo, size_t hi) -> int {
        const size_t g = (size_t)(&d - &ctx->devs[0]);
        const uint8_t *ddst; uint32_t dlen;
        if (int rc = stage_dst(d, dst, dst_len, true, &ddst, &dlen)) return rc;
        const size_t chunk = ctx->verify_chunk(d);
        Pipe pipe(d);
        for (size_t c = lo; c < hi; c += chunk) {
            size_t e = std::min(hi, c + chunk), m = e - c;
            if (int rc = pipe.begin(pin_ok ? 0 : pad64(m))) return rc;
            const uint8_t *dm; const uint64_t *dof; uint64_t base;
            if (int rc = stage_msgs(d, &pipe, msgs, offsets, c, e, &dm, &dof, &base)) return rc;
            uint8_t *dsig, *dok; uint32_t *dki;
            if (int rc = ensure_t(d, stage_buf(S_IN_A, pipe.set()), m, &dki)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_IN_B, pipe.set()), m * 64, &dsig)) return rc;
            if (int rc = ensure_t(d, stage_buf(S_OUT, pipe.set()), m, &dok)) return rc;
            if (int rc = pipe.h2d(dki, key_index + c, m * sizeof(uint32_t))) return rc;
            if (int rc = pipe.h2d(dsig, sigs + c * 64, m * 64)) return rc;
            if (int rc = pipe.inputs_ready()) return rc;
            if (int rc = dev_verify_prepared(d, (const line_t<fp2> *)pr->tab[g], (const uint8_t *)pr->fl[g], (const g2m *)pr->pts[g], dki, dm, dof, base, dsig, m, ddst, dlen, flags, dok))
                return rc;
            if (int rc = pipe.d2h(ok + c, pin_ok, dok, m)) return rc;
            if (int rc = pipe.end()) return rc;
        }
        return pipe.finish();
    });
}

namespace {
// Every device of the ctx reduces its contiguous slice of the (pk_i, m_i) pairs to ONE record in its NCCL send buffer
// (device 0 -- rank 0 of a one-process-per-GPU job -- also takes the aggregate signature), the records are all-gathered
// over NVLink (one ncclAllGather per device inside one NCCL group), and the devices end with every rank's record in
// S_AGG_RECV.  Host threads: one per device for the shard, the caller's thread for the collective.
int aggregate_shards_gathered(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                              const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags) {
    size_t rpk = variant == BN254_MIN_SIG ? 128 : 64, rsig = variant == BN254_MIN_SIG ? 64 : 128;
    size_t nd = ctx->devs.size();
    static const uint64_t zero_off[1] = {0};
    const uint64_t *offs = n ? offsets : zero_off;
    std::vector<std::thread> th;
    for (size_t g = 0; g < nd; g++) {
        size_t lo = n * g / nd, hi = n * (g + 1) / nd;
        Dev *d = &ctx->devs[g];
        d->rc = 0;
        auto work = [=]() {
            cudaSetDevice(d->id);
            d->rc = [&]() -> int {
                Dev &dd = *d;
                size_t m = hi - lo;
                const uint8_t *ddst; uint32_t dlen;
                if (int rc = stage_dst(dd, dst, dst_len, true, &ddst, &dlen)) return rc;
                const uint8_t *dm; const uint64_t *dof; uint64_t base;
                if (int rc = stage_msgs(dd, nullptr, msgs, offs, lo, hi, &dm, &dof, &base)) return rc;
                uint8_t *dpk, *dsig = nullptr;
                agg_rec *send;
                if (int rc = ensure_t(dd, S_IN_A, m * rpk + 16, &dpk)) return rc;
                if (int rc = ensure_t(dd, S_AGG_SEND, 1, &send)) return rc;
                if (m) CK(dd, cudaMemcpyAsync(dpk, pks + lo * rpk, m * rpk, cudaMemcpyHostToDevice, dd.st));
                if (g == 0 && agg_sig) {
                    if (int rc = ensure_t(dd, S_IN_B, rsig, &dsig)) return rc;
                    CK(dd, cudaMemcpyAsync(dsig, agg_sig, rsig, cudaMemcpyHostToDevice, dd.st));
                }
                return dev_aggregate_rec(dd, variant, dpk, dm, dof, base, m, dsig, ddst, dlen, flags, send);
            }();
        };
        if (nd == 1) work();
        else th.emplace_back(work);
    }
    for (auto &t : th) t.join();
    for (auto &d : ctx->devs)
        if (d.rc) {
            ctx->err = d.err;
            return d.rc;
        }
    // the exchange: every device's all-gather inside one NCCL group (a lone device without a communicator copies)
    const bool grouped = nd > 1;
    if (grouped) {
        if (int rc = need_comm(ctx->devs[0])) {
            ctx->err = ctx->devs[0].err;
            return rc;
        }
        bn_nccl::api().GroupStart();
    }
    int rc = 0;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.id);
        agg_rec *recv;
        rc = dev_gather_recs(d, (const agg_rec *)d.bufs[S_AGG_SEND].p, &recv);
        if (rc) {
            ctx->err = d.err;
            break;
        }
    }
    if (grouped) {
        int r = bn_nccl::api().GroupEnd();
        if (!rc && r != bn_nccl::kSuccess) {
            ctx->err = std::string("ncclGroupEnd: ") + bn_nccl::api().GetErrorString(r);
            rc = BN254_ERR_NCCL_BASE - r;
        }
    }
    return rc;
}
}  // namespace

int bn254_aggregate_partial(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                            const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *partial, uint8_t *bad) {
    if (!ctx) return BN254_ERR_ARG;
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (!partial || !bad || (n && (!pks || !offsets))) return fail_arg(ctx, "null pointer");
    if (n)
        if (int rc = check_offsets(ctx, offsets, n)) return rc;
    // A multi-device ctx is ONE shard towards its caller: its devices' records are gathered (NCCL) and device 0 multiplies them.
    // A one-device ctx must not touch its job-level communicator here (the other ranks are not in this call).
    Dev &d = ctx->devs[0];
    const int keep_ranks = d.nranks;
    bn_nccl::comm_t keep_comm = d.comm;
    if (ctx->devs.size() == 1) { d.nranks = 1; d.comm = nullptr; }
    int rc = aggregate_shards_gathered(ctx, variant, pks, msgs, offsets, n, agg_sig, dst, dst_len, flags);
    const int k = d.nranks;
    d.nranks = keep_ranks;
    d.comm = keep_comm;
    if (rc) return rc;
    cudaSetDevice(d.id);
    auto body = [&]() -> int {
        fp12 *prod; int *dbad; uint8_t *dout;
        if (int rc2 = ensure_t(d, S_F2, 4, &prod)) return rc2;
        if (int rc2 = ensure_t(d, S_BAD, 4, &dbad)) return rc2;
        if (int rc2 = ensure_t(d, S_OUT, 384 + 16, &dout)) return rc2;
        CK(d, cudaMemsetAsync(dbad, 0, sizeof(int), d.st));
        LAUNCH(d, k_gather_product_l2, 32, 32, prod, dbad, (const agg_rec *)d.bufs[S_AGG_RECV].p, k);
        LAUNCH(d, k_fp12_encode, 1, BN_TPB_LIGHT, dout, (uint8_t *)nullptr, prod, (const uint8_t *)nullptr, (size_t)1);
        int hbad = 0;
        CK(d, cudaMemcpyAsync(partial, dout, 384, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaMemcpyAsync(&hbad, dbad, sizeof(int), cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        *bad = hbad ? 1 : 0;
        return 0;
    };
    rc = body();
    if (rc) ctx->err = d.err;
    if (!rc) return bn254_sync(ctx);   // the other devices' gathers
    return rc;
}

int bn254_aggregate_finish(bn254_ctx *ctx, const uint8_t *partials, const uint8_t *bad, size_t kparts, uint8_t *ok) {
    if (!ctx || !partials || !ok || kparts == 0) return fail_arg(ctx, "null pointer / no partials");
    Dev &d = ctx->devs[0];
    d.rc = 0;
    cudaSetDevice(d.id);
    uint8_t anybad = 0;
    if (bad)
        for (size_t i = 0; i < kparts; i++) anybad |= bad[i];
    auto body = [&]() -> int {
        uint8_t *din, *dok, *dbadz;
        if (int rc = ensure_t(d, S_IN_A, kparts * 384, &din)) return rc;
        if (int rc = ensure_t(d, S_OUT, 16, &dok)) return rc;
        if (int rc = ensure_t(d, S_STATUS, kparts + 16, &dbadz)) return rc;
        CK(d, cudaMemcpyAsync(din, partials, kparts * 384, cudaMemcpyHostToDevice, d.st));
        CK(d, cudaMemsetAsync(dbadz, 0, kparts, d.st));
        if (int rc = dev_aggregate_finish(d, din, dbadz, kparts, dok)) return rc;
        uint8_t h = 0;
        CK(d, cudaMemcpyAsync(&h, dok, 1, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        *ok = (h && !anybad) ? 1 : 0;
        return 0;
    };
    int rc = body();
    if (rc) ctx->err = d.err;
    return rc;
}

// One call = the whole predicate on this ctx's devices: shards -> all-gather -> ONE post-gather kernel on device 0.
int bn254_aggregate_verify(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                           const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    if (!ctx || !agg_sig || !ok) return fail_arg(ctx, "null pointer");
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (n && (!pks || !offsets)) return fail_arg(ctx, "null pointer");
    if (n)
        if (int rc = check_offsets(ctx, offsets, n)) return rc;
    Dev &d = ctx->devs[0];
    const int keep_ranks = d.nranks;
    bn_nccl::comm_t keep_comm = d.comm;
    if (ctx->devs.size() == 1) { d.nranks = 1; d.comm = nullptr; }   // a lone-device ctx verifies the whole aggregate itself
    int rc = aggregate_shards_gathered(ctx, variant, pks, msgs, offsets, n, agg_sig, dst, dst_len, flags);
    auto body = [&]() -> int {
        uint8_t *dok;
        if (int rc2 = ensure_t(d, S_OUT, 16, &dok)) return rc2;
        if (int rc2 = dev_gather_finish(d, (const agg_rec *)d.bufs[S_AGG_RECV].p, dok)) return rc2;
        uint8_t h = 0;
        CK(d, cudaMemcpyAsync(&h, dok, 1, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        *ok = h;
        return 0;
    };
    if (!rc) {
        cudaSetDevice(d.id);
        rc = body();
        if (rc) ctx->err = d.err;
    }
    d.nranks = keep_ranks;
    d.comm = keep_comm;
    if (!rc) return bn254_sync(ctx);
    return rc;
}

// The same predicate for a job of one process per GPU (torchrun / MPI): EVERY rank calls this with ITS shard of the
// (pk_i, m_i) pairs; agg_sig is read on rank 0 only (may be NULL elsewhere).  One ncclAllGather of 400 bytes per rank on
// the library stream, then every rank runs the post-gather kernel, so *ok is the job's verdict on every rank.
int bn254_aggregate_verify_dist(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    if (!ctx || !ok) return fail_arg(ctx, "null pointer");
    if (ctx->devs.size() != 1) return fail_arg(ctx, "bn254_aggregate_verify_dist needs a one-device ctx");
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (n && (!pks || !offsets)) return fail_arg(ctx, "null pointer");
    Dev &d = ctx->devs[0];
    if (d.rank == 0 && !agg_sig) return fail_arg(ctx, "rank 0 needs the aggregate signature");
    if (n)
        if (int rc = check_offsets(ctx, offsets, n)) return rc;
    int rc = aggregate_shards_gathered(ctx, variant, pks, msgs, offsets, n, d.rank == 0 ? agg_sig : nullptr, dst, dst_len, flags);
    if (rc) return rc;
    cudaSetDevice(d.id);
    auto body = [&]() -> int {
        uint8_t *dok;
        if (int rc2 = ensure_t(d, S_OUT, 16, &dok)) return rc2;
        if (int rc2 = dev_gather_finish(d, (const agg_rec *)d.bufs[S_AGG_RECV].p, dok)) return rc2;
        uint8_t h = 0;
        CK(d, cudaMemcpyAsync(&h, dok, 1, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        *ok = h;
        return 0;
    };
    rc = body();
    if (rc) ctx->err = d.err;
    return rc;
}

// ---- codecs / predicates / scalar multiplication ---------------------------------------------------
int bn254_g1_compress(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) {
    return host_map(ctx, n, pts, 64, nullptr, 0, out, 32, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *) {
        g1m *P; uint8_t *fl;
        if (int rc = ensure_t(d, S_P0, m, &P)) return rc;
        if (int rc = ensure_t(d, S_FLP0, m, &fl)) return rc;
        LAUNCH(d, k_g1_decode, m, BN_TPB_LIGHT, P, fl, a, m, 0);
        LAUNCH(d, k_g1_compress, m, BN_TPB_LIGHT, o, P, fl, m);
        return 0;
    });
}
int bn254_g2_compress(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) {
    return host_map(ctx, n, pts, 128, nullptr, 0, out, 64, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *) {
        g2m *P; uint8_t *fl;
        if (int rc = ensure_t(d, S_Q0, m, &P)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, m, &fl)) return rc;
        LAUNCH(d, k_g2_decode, m, BN_TPB_LIGHT, P, fl, a, m, 0);
        LAUNCH(d, k_g2_compress, m, BN_TPB_LIGHT, o, P, fl, m);
        return 0;
    });
}
int bn254_g1_decompress(bn254_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, uint8_t *status) {
    return host_map(ctx, n, in, 32, nullptr, 0, out, 64, status, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *s) {
        g1m *P; uint8_t *fl;
        if (int rc = ensure_t(d, S_P0, m, &P)) return rc;
        if (int rc = ensure_t(d, S_FLP0, m, &fl)) return rc;
        LAUNCH(d, k_g1_decompress, m, BN_TPB_HEAVY, P, fl, a, m);
        LAUNCH(d, k_g1_encode, m, BN_TPB_LIGHT, o, s, P, fl, m);
        return 0;
    });
}
int bn254_g2_decompress(bn254_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, uint8_t *status) {
    return host_map(ctx, n, in, 64, nullptr, 0, out, 128, status, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *s) {
        g2m *P; uint8_t *fl;
        if (int rc = ensure_t(d, S_Q0, m, &P)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, m, &fl)) return rc;
        LAUNCH(d, k_g2_decompress, m, BN_TPB_HEAVY, P, fl, a, m);
        LAUNCH(d, k_g2_encode, m, BN_TPB_LIGHT, o, s, P, fl, m);
        return 0;
    });
}
int bn254_g1_is_on_curve(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) {
    return host_map(ctx, n, pts, 64, nullptr, 0, out, 1, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *) {
        LAUNCH(d, k_g1_on_curve, m, BN_TPB_LIGHT, o, a, m);
        return 0;
    });
}
int bn254_g2_is_on_curve(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) {
    return host_map(ctx, n, pts, 128, nullptr, 0, out, 1, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *) {
        LAUNCH(d, k_g2_on_curve, m, BN_TPB_LIGHT, o, a, m);
        return 0;
    });
}
int bn254_g1_is_torsion_free(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) {
    return host_map(ctx, n, pts, 64, nullptr, 0, out, 1, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *) {
        LAUNCH(d, k_g1_torsion_free, m, BN_TPB_HEAVY, o, a, m);
        return 0;
    });
}
int bn254_g2_is_torsion_free(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) {
    return host_map(ctx, n, pts, 128, nullptr, 0, out, 1, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *, size_t m, uint8_t *o, uint8_t *) {
        LAUNCH(d, k_g2_torsion_free, m, BN_TPB_HEAVY, o, a, m);
        return 0;
    });
}
int bn254_g1_mul(bn254_ctx *ctx, const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out) {
    if (n && !scalars) return fail_arg(ctx, "null pointer");
    return host_map(ctx, n, pts, 64, scalars, 32, out, 64, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *b, size_t m, uint8_t *o, uint8_t *) {
        g1m *P, *R; uint8_t *fl, *flo;
        if (int rc = ensure_t(d, S_P0, m, &P)) return rc;
        if (int rc = ensure_t(d, S_P1, m, &R)) return rc;
        if (int rc = ensure_t(d, S_FLP0, m, &fl)) return rc;
        if (int rc = ensure_t(d, S_FLP1, m, &flo)) return rc;
        LAUNCH(d, k_g1_decode, m, BN_TPB_LIGHT, P, fl, a, m, 0);
        LAUNCH(d, k_g1_mul, m, BN_TPB_HEAVY, R, flo, P, fl, b, m);
        LAUNCH(d, k_g1_encode, m, BN_TPB_LIGHT, o, (uint8_t *)nullptr, R, flo, m);
        return 0;
    });
}
int bn254_g2_mul(bn254_ctx *ctx, const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out) {
    if (n && !scalars) return fail_arg(ctx, "null pointer");
    return host_map(ctx, n, pts, 128, scalars, 32, out, 128, nullptr, [](Dev &d, const uint8_t *a, const uint8_t *b, size_t m, uint8_t *o, uint8_t *) {
        g2m *P, *R; uint8_t *fl, *flo;
        if (int rc = ensure_t(d, S_Q0, m, &P)) return rc;
        if (int rc = ensure_t(d, S_Q1, m, &R)) return rc;
        if (int rc = ensure_t(d, S_FLQ0, m, &fl)) return rc;
        if (int rc = ensure_t(d, S_FLQ1, m, &flo)) return rc;
        LAUNCH(d, k_g2_decode, m, BN_TPB_LIGHT, P, fl, a, m, 0);
        LAUNCH(d, k_g2_mul, m, BN_TPB_HEAVY, R, flo, P, fl, b, m);
        LAUNCH(d, k_g2_encode, m, BN_TPB_LIGHT, o, (uint8_t *)nullptr, R, flo, m);
        return 0;
    });
}

static int sum_impl(bn254_ctx *ctx, int group, const uint8_t *pts, size_t n, uint8_t *out) {
    if (!ctx || !out || (n && !pts)) return fail_arg(ctx, "null pointer");
    Dev &d = ctx->devs[0];
    d.rc = 0;
    cudaSetDevice(d.id);
    size_t rec = group == 1 ? 64 : 128;
    int bad = 0;
    auto body = [&]() -> int {
        uint8_t *din, *dout, *hfl = nullptr;
        if (int rc = ensure_t(d, S_IN_A, n * rec + 16, &din)) return rc;
        if (int rc = ensure_t(d, S_OUT, rec + 16, &dout)) return rc;
        if (n) CK(d, cudaMemcpyAsync(din, pts, n * rec, cudaMemcpyHostToDevice, d.st));
        size_t m0 = std::max<size_t>(1, (n + 15) / 16);
        if (group == 1) {
            g1m *P, *A; uint8_t *fl, *fla;
            if (int rc = ensure_t(d, S_P0, n + 1, &P)) return rc;
            if (int rc = ensure_t(d, S_FLP0, n + 1, &fl)) return rc;
            if (int rc = ensure_t(d, S_P1, 2 * m0 + 2, &A)) return rc;
            if (int rc = ensure_t(d, S_FLP1, 2 * m0 + 2, &fla)) return rc;
            LAUNCH(d, k_g1_decode, n, BN_TPB_LIGHT, P, fl, din, n, 0);
            const g1m *src = P; const uint8_t *sfl = fl; size_t cur = n; g1m *dp = A; uint8_t *dfl = fla;
            while (true) {
                size_t m = std::max<size_t>(1, (cur + 15) / 16);
                LAUNCH(d, k_g1_sum, m, BN_TPB_HEAVY, dp, dfl, src, sfl, cur, m);
                if (m == 1) break;
                src = dp; sfl = dfl; cur = m;
                dp = (dp == A) ? A + m0 : A; dfl = (dfl == fla) ? fla + m0 : fla;
            }
            LAUNCH(d, k_g1_encode, 1, BN_TPB_LIGHT, dout, dout + rec, dp, dfl, (size_t)1);
            hfl = dout + rec;
        } else {
            g2m *P, *A; uint8_t *fl, *fla;
            if (int rc = ensure_t(d, S_Q0, n + 1, &P)) return rc;
            if (int rc = ensure_t(d, S_FLQ0, n + 1, &fl)) return rc;
            if (int rc = ensure_t(d, S_Q1, 2 * m0 + 2, &A)) return rc;
            if (int rc = ensure_t(d, S_FLQ1, 2 * m0 + 2, &fla)) return rc;
            LAUNCH(d, k_g2_decode, n, BN_TPB_LIGHT, P, fl, din, n, 0);
            const g2m *src = P; const uint8_t *sfl = fl; size_t cur = n; g2m *dp = A; uint8_t *dfl = fla;
            while (true) {
                size_t m = std::max<size_t>(1, (cur + 15) / 16);
                LAUNCH(d, k_g2_sum, m, BN_TPB_HEAVY, dp, dfl, src, sfl, cur, m);
                if (m == 1) break;
                src = dp; sfl = dfl; cur = m;
                dp = (dp == A) ? A + m0 : A; dfl = (dfl == fla) ? fla + m0 : fla;
            }
            LAUNCH(d, k_g2_encode, 1, BN_TPB_LIGHT, dout, dout + rec, dp, dfl, (size_t)1);
            hfl = dout + rec;
        }
        uint8_t st = 1;
        CK(d, cudaMemcpyAsync(out, dout, rec, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaMemcpyAsync(&st, hfl, 1, cudaMemcpyDeviceToHost, d.st));
        CK(d, cudaStreamSynchronize(d.st));
        bad = !st;
        return 0;
    };
    int rc = body();
    if (rc) {
        ctx->err = d.err;
        return rc;
    }
    if (bad) return group == 1 ? BN254_ERR_INVALID_G1_BYTES : BN254_ERR_INVALID_G2_BYTES;
    return 0;
}
int bn254_g1_sum(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) { return sum_impl(ctx, 1, pts, n, out); }
int bn254_g2_sum(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out) { return sum_impl(ctx, 2, pts, n, out); }

// ---- device-resident entry points ------------------------------------------------------------------
#define DEV_PROLOGUE()                \
    if (!ctx) return BN254_ERR_ARG;   \
    Dev &d = ctx->devs[0];            \
    d.rc = 0;                         \
    cudaSetDevice(d.id);
#define DEV_EPILOGUE(rc)              \
    if (rc) ctx->err = d.err;         \
    return rc;

int bn254_hash_to_g1_dev(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, int count,
                         uint8_t *out, uint8_t *status) {
    DEV_PROLOGUE();
    if (count != 1 && count != 2) return fail_arg(ctx, "count must be 1 (encode) or 2 (hash)");
    const uint8_t *ddst; uint32_t dlen;
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = dev_hash_wire(d, 1, msgs, offsets, 0, n, ddst, dlen, count, out, status);
    DEV_EPILOGUE(rc);
}
int bn254_hash_to_g2_dev(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, int count,
                         uint8_t *out, uint8_t *status) {
    DEV_PROLOGUE();
    if (count != 1 && count != 2) return fail_arg(ctx, "count must be 1 (encode) or 2 (hash)");
    const uint8_t *ddst; uint32_t dlen;
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = dev_hash_wire(d, 2, msgs, offsets, 0, n, ddst, dlen, count, out, status);
    DEV_EPILOGUE(rc);
}
int bn254_pairing_batch_dev(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *gt, uint8_t *status) {
    DEV_PROLOGUE();
    int rc = dev_pairing(d, g1, g2, n, gt, status, 0);
    DEV_EPILOGUE(rc);
}
int bn254_verify_batch_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, const uint8_t *sigs,
                           size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    DEV_PROLOGUE();
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    const uint8_t *ddst; uint32_t dlen;
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = dev_verify(d, variant, pks, msgs, offsets, 0, sigs, n, ddst, dlen, flags, ok);
    DEV_EPILOGUE(rc);
}
int bn254_verify_batch_prepared_dev(bn254_ctx *ctx, const bn254_g2_prepared *pr, const uint32_t *key_index, const uint8_t *msgs, const uint64_t *offsets,
                                    const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    DEV_PROLOGUE();
    if (!pr || pr->ctx != ctx) return fail_arg(ctx, "prepared keys belong to another ctx");
    const uint8_t *ddst; uint32_t dlen;
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = dev_verify_prepared(d, (const line_t<fp2> *)pr->tab[0], (const uint8_t *)pr->fl[0], (const g2m *)pr->pts[0], key_index, msgs, offsets, 0, sigs, n, ddst, dlen, flags, ok);
    DEV_EPILOGUE(rc);
}
int bn254_verify_batch_randomized_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets,
                                      const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, const uint8_t *seed,
                                      int rand_bits, size_t group, uint8_t *ok_groups) {
    DEV_PROLOGUE();
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (rand_bits != 64 && rand_bits != 128) return fail_arg(ctx, "rand_bits must be 64 or 128");
    if (group == 0 || group > BN254_RBV_MAX_GROUP) return fail_arg(ctx, "group must be in 1..BN254_RBV_MAX_GROUP");
    if (n == 0) return 0;
    const uint8_t *ddst; uint32_t dlen;
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = dev_verify_rand(d, variant, pks, msgs, offsets, 0, sigs, n, group, seed, 0, rand_bits, ddst, dlen, flags, ok_groups);
    DEV_EPILOGUE(rc);
}
// device pointers (device 0 of the ctx) for everything but key_offsets, which stays a HOST array: the reduction plan is built
// from it.  One chunk: every scratch array is sized for all n tuples.
int bn254_verify_same_key_randomized_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint64_t *key_offsets, size_t n_keys,
                                         const uint8_t *msgs, const uint64_t *offsets, const uint8_t *sigs, const uint8_t *dst, size_t dst_len,
                                         uint32_t flags, const uint8_t *seed, int rand_bits, uint8_t *ok_keys) {
    DEV_PROLOGUE();
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (rand_bits != 64 && rand_bits != 128) return fail_arg(ctx, "rand_bits must be 64 or 128");
    if (n_keys == 0) return 0;
    if (!pks || !key_offsets || !seed || !ok_keys) return fail_arg(ctx, "null pointer");
    if (int rc = check_offsets(ctx, key_offsets, n_keys)) return rc;
    if (key_offsets[0] != 0) return fail_arg(ctx, "key_offsets[0] must be 0 for device buffers");
    const size_t n = (size_t)key_offsets[n_keys];
    if (n && (!offsets || !sigs)) return fail_arg(ctx, "null pointer");
    const uint8_t *ddst; uint32_t dlen;
    uint64_t *dplan = nullptr;
    SkPlan plan;
    sk_plan(plan, key_offsets, n_keys, !sk_use_msm(d, (size_t)key_offsets[n_keys], n_keys));
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = ensure_t(d, S_IN_C, plan.blob.size(), &dplan);
    if (!rc && cudaMemcpyAsync(dplan, plan.blob.data(), plan.blob.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, d.st) != cudaSuccess) {
        d.err = "cudaMemcpyAsync of the reduction plan failed";
        rc = BN254_ERR_CUDA_BASE;
    }
    if (!rc) rc = dev_verify_same_key(d, variant, pks, n_keys, dplan, plan, msgs, offsets, 0, sigs, n, seed, 0, rand_bits, ddst, dlen, flags, ok_keys);
    DEV_EPILOGUE(rc);
}
int bn254_aggregate_partial_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *partial, uint8_t *bad) {
    DEV_PROLOGUE();
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    const uint8_t *ddst; uint32_t dlen;
    int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen);
    if (!rc) rc = dev_aggregate_partial(d, variant, pks, msgs, offsets, 0, n, agg_sig, ddst, dlen, flags, partial, bad);
    DEV_EPILOGUE(rc);
}

// device-resident form of bn254_aggregate_verify_dist: everything (shard, all-gather, post-gather kernel) is enqueued on
// bn254_stream(ctx); *ok (device byte) holds the job's verdict on every rank once the stream has drained
int bn254_aggregate_verify_dist_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                    const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok) {
    DEV_PROLOGUE();
    if (ctx->devs.size() != 1) return fail_arg(ctx, "needs a one-device ctx");
    if (variant != BN254_MIN_SIG && variant != BN254_MIN_PK) return fail_arg(ctx, "variant must be 0 (min-sig) or 1 (min-pk)");
    if (d.rank == 0 && !agg_sig) return fail_arg(ctx, "rank 0 needs the aggregate signature");
    auto body = [&]() -> int {
        const uint8_t *ddst; uint32_t dlen;
        agg_rec *send, *recv;
        if (int rc = stage_dst(d, dst, dst_len, false, &ddst, &dlen)) return rc;
        if (int rc = ensure_t(d, S_AGG_SEND, 1, &send)) return rc;
        if (int rc = dev_aggregate_rec(d, variant, pks, msgs, offsets, 0, n, d.rank == 0 ? agg_sig : nullptr, ddst, dlen, flags, send)) return rc;
        if (int rc = dev_gather_recs(d, send, &recv)) return rc;
        return dev_gather_finish(d, recv, ok);
    };
    int rc = body();
    DEV_EPILOGUE(rc);
}

int bn254_imad_peak(bn254_ctx *ctx, int mode, int iters, double *ops_per_s, double *ms_out) {
    DEV_PROLOGUE();
    if (iters < 1 || !ops_per_s) return fail_arg(ctx, "bad argument");
    auto body = [&]() -> int {
        cudaDeviceProp prop;
        CK(d, cudaGetDeviceProperties(&prop, d.id));
        uint32_t *sink;
        if (int rc = ensure_t(d, S_TMP, 16, &sink)) return rc;
        int nblk = prop.multiProcessorCount * 8;  // 8 x 256 threads per SM = full occupancy
        cudaEvent_t e0, e1;
        CK(d, cudaEventCreate(&e0));
        CK(d, cudaEventCreate(&e1));
        k_imad_peak<<<nblk, 256, 0, d.st>>>(sink, iters / 8 + 1, mode);  // warm-up
        d.launches++;
        CK(d, cudaEventRecord(e0, d.st));
        k_imad_peak<<<nblk, 256, 0, d.st>>>(sink, iters, mode);
        d.launches++;
        CK(d, cudaEventRecord(e1, d.st));
        CK(d, cudaEventSynchronize(e1));
        CK(d, cudaGetLastError());
        float ms = 0;
        CK(d, cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        double ops = (double)nblk * 256.0 * (double)iters * 64.0;
        *ops_per_s = ops / (ms * 1e-3);
        if (ms_out) *ms_out = ms;
        return 0;
    };
    int rc = body();
    DEV_EPILOGUE(rc);
}

// Field-level micro-benchmark (k_fp_peak): warps_per_sm warps of dependent fp_dot2 (mode 0) / fp_mul (mode 1)
// chains per SM; returns canonical IMAD-ops/s (264 per product, 400 + 16 per fused dot product).
int bn254_fp_peak(bn254_ctx *ctx, int mode, int warps_per_sm, int iters, double *ops_per_s) {
    DEV_PROLOGUE();
    if (iters < 1 || warps_per_sm < 1 || !ops_per_s) return fail_arg(ctx, "bad argument");
    auto body = [&]() -> int {
        cudaDeviceProp prop;
        CK(d, cudaGetDeviceProperties(&prop, d.id));
        uint32_t *sink;
        if (int rc = ensure_t(d, S_TMP, 16, &sink)) return rc;
        int nblk = prop.multiProcessorCount * warps_per_sm;
        cudaEvent_t e0, e1;
        CK(d, cudaEventCreate(&e0));
        CK(d, cudaEventCreate(&e1));
        k_fp_peak<<<nblk, 32, 0, d.st>>>(sink, iters / 8 + 1, mode);
        d.launches++;
        CK(d, cudaEventRecord(e0, d.st));
        k_fp_peak<<<nblk, 32, 0, d.st>>>(sink, iters, mode);
        d.launches++;
        CK(d, cudaEventRecord(e1, d.st));
        CK(d, cudaEventSynchronize(e1));
        CK(d, cudaGetLastError());
        float ms = 0;
        CK(d, cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        double per = mode == 0 ? 2.0 * 264.0 : (mode == 1 ? 264.0 : 4.0 * 264.0);   // a dot product does the work of two products
        *ops_per_s = (double)nblk * 32.0 * (double)iters * per / (ms * 1e-3);
        return 0;
    };
    int rc = body();
    DEV_EPILOGUE(rc);
}

}  // extern "C"
