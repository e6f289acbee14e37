// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the device headers (bls_bn254_b200/csrc/*.cuh)
// with g++ so the exact per-item bodies the CUDA kernels run can be unit-tested on a machine without
// a GPU (carry chains are emulated in C, see fp.cuh).  Never linked into libbls_bn254_b200.so and
// never reachable from the C ABI: the product has no CPU path.
#include <atomic>
#include <cstring>

#include "../../bls_bn254_b200/csrc/items.cuh"

using namespace bn;
#define EXPORT extern "C" __attribute__((visibility("default")))

// executed-work counter (fp.cuh: BN_COUNT_OPS): INT32 multiply-add operations issued by the multiplier routines since
// the last reset, summed over all emulated lanes.  tools/count_ops.py freezes the per-item figures.
static std::atomic<unsigned long long> g_ops{0};
extern "C" void bn_host_count_ops(unsigned ops) { g_ops.fetch_add(ops, std::memory_order_relaxed); }
EXPORT unsigned long long sim_ops(void) { return g_ops.load(); }
EXPORT void sim_ops_reset(void) { g_ops.store(0); }

EXPORT void sim_fp_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    fp x, y, z;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    fp_mul(z, x, y);
    memcpy(r, z.v, 32);
}
EXPORT void sim_fp_add(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    fp x, y, z;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    fp_add(z, x, y);
    memcpy(r, z.v, 32);
}
EXPORT void sim_fp_sub(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    fp x, y, z;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    fp_sub(z, x, y);
    memcpy(r, z.v, 32);
}
// canonical BE in / out
EXPORT int sim_fp2_sqrt(uint8_t *out, const uint8_t *in) {
    fp2 a, r;
    fp_from_be(a.c0, in);
    fp_from_be(a.c1, in + 32);
    bool ok = fp2_sqrt(r, a);
    fp_to_be(out, r.c0);
    fp_to_be(out + 32, r.c1);
    return ok;
}
EXPORT void sim_hash_to_g1(const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, int count, uint8_t *out,
                           uint8_t *status) {
    for (size_t i = 0; i < n; i++) {
        g1m p;
        uint8_t fl;
        item_hash_g1(p, fl, msgs + off[i], off[i + 1] - off[i], dst, (uint32_t)dst_len, count);
        item_g1_encode(out + 64 * i, p, fl);
        if (status) status[i] = !(fl & FL_BAD);
    }
}
EXPORT void sim_hash_to_g2(const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, int count, uint8_t *out,
                           uint8_t *status) {
    for (size_t i = 0; i < n; i++) {
        g2m p;
        uint8_t fl;
        item_hash_g2(p, fl, msgs + off[i], off[i + 1] - off[i], dst, (uint32_t)dst_len, count);
        item_g2_encode(out + 128 * i, p, fl);
        if (status) status[i] = !(fl & FL_BAD);
    }
}
EXPORT void sim_expand_message_xmd(const uint8_t *msg, size_t msg_len, const uint8_t *dst, size_t dst_len, uint8_t *out, size_t n) {
    uint8_t d2[32];
    if (dst_len > 255) {
        oversize_dst(d2, dst, dst_len);
        dst = d2;
        dst_len = 32;
    }
    expand_message_xmd(out, (uint32_t)n, msg, msg_len, dst, (uint32_t)dst_len);
}
// mode 0: pairing, 1: miller loop only
EXPORT void sim_pairing_batch(const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *out, uint8_t *status, int mode) {
    for (size_t i = 0; i < n; i++) {
        g1m p;
        g2m q;
        uint8_t flp, flq, fl;
        fp12 f;
        item_g1_decode(p, flp, g1 + 64 * i, false);
        item_g2_decode(q, flq, g2 + 128 * i, false);
        item_miller1(f, fl, p, flp, q, flq);
        if (mode == 0) item_final_exp(f);
        if (fl & FL_BAD) memset(out + 384 * i, 0, 384);
        else fp12_to_bytes(out + 384 * i, f);
        if (status) status[i] = !(fl & FL_BAD);
    }
}
EXPORT void sim_final_exp(const uint8_t *in, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) {
        fp12 f;
        fp12_from_bytes(f, in + 384 * i);
        item_final_exp(f);
        fp12_to_bytes(out + 384 * i, f);
    }
}
// 2-pair Miller loop: fixed != 0 -> pair 0 is (p0, -g2) via the line table (q0 ignored)
EXPORT void sim_miller2(const uint8_t *p0, const uint8_t *q0, const uint8_t *p1, const uint8_t *q1, int fixed, uint8_t *out) {
    g1m a, c;
    g2m b, d;
    uint8_t fa, fb = 0, fc, fd, fl;
    fp12 f;
    item_g1_decode(a, fa, p0, false);
    if (!fixed) item_g2_decode(b, fb, q0, false);
    item_g1_decode(c, fc, p1, false);
    item_g2_decode(d, fd, q1, false);
    if (fixed) item_miller2<true>(f, fl, a, fa, nullptr, 0, c, fc, d, fd);
    else item_miller2<false>(f, fl, a, fa, &b, fb, c, fc, d, fd);
    fp12_to_bytes(out, f);
}
EXPORT void sim_g1_mul(const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) {
        g1m p, r;
        uint8_t fl, flo;
        item_g1_decode(p, fl, pts + 64 * i, false);
        item_g1_mul(r, flo, p, fl, sc + 32 * i);
        item_g1_encode(out + 64 * i, r, flo);
    }
}
EXPORT void sim_g2_mul(const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) {
        g2m p, r;
        uint8_t fl, flo;
        item_g2_decode(p, fl, pts + 128 * i, false);
        item_g2_mul(r, flo, p, fl, sc + 32 * i);
        item_g2_encode(out + 128 * i, r, flo);
    }
}
// Fp12 tower operations of tower.cuh one at a time (384-byte wire elements, pairings.rs:499-514), for both Fp2 representations
// where that applies.  op: 0 mul, 1 sqr, 2 inv, 3 conj, 4/5/6 frobenius p^1 / p^2 / p^3, 7 cyclotomic square (a must be in the
// cyclotomic subgroup), 8 sparse product a * (c0 + c3 w + c4 w^3) with (c0, c3, c4) = the first three Fp2 of b, 9 the paired
// sparse product a * l * l' of the two-pair Miller loops with l from the first and l' from the last three Fp2 of b
EXPORT int sim_fp12_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    fp12 x, y, r;
    if (!fp12_from_bytes(x, a)) return -1;
    if (b && !fp12_from_bytes(y, b)) return -1;
    switch (op) {
        case 0: fp12_mul(r, x, y); break;
        case 1: fp12_sqr(r, x); break;
        case 2: fp12_inv(r, x); break;
        case 3: fp12_conj(r, x); break;
        case 4: case 5: case 6: fp12_frob(r, x, op - 3); break;
        case 7: cyclotomic_square(r, x); break;
        case 8: r = x; fp12_mul_by_034(r, y.c0.c0, y.c0.c1, y.c0.c2); break;
        case 9: r = x; fp12_mul_by_034_034(r, y.c0.c0, y.c0.c1, y.c0.c2, y.c1.c0, y.c1.c1, y.c1.c2); break;
        default: return -2;
    }
    fp12_to_bytes(out, r);
    return 0;
}
// quadratic-residue test: binary Jacobi symbol vs the Euler exponentiation it replaces (raw limbs, < p)
EXPORT void sim_fp_is_square(const uint32_t *a, size_t n, uint8_t *jacobi, uint8_t *euler) {
    for (size_t i = 0; i < n; i++) {
        fp x;
        for (int j = 0; j < 8; j++) x.v[j] = a[8 * i + j];
        jacobi[i] = fp_is_square_jacobi(x) ? 1 : 0;
        euler[i] = fp_is_square_euler(x) ? 1 : 0;
    }
}
// modular inversion: binary extended GCD vs the Fermat exponentiation it replaces (raw Montgomery limbs, < p)
EXPORT void sim_fp_inv(const uint32_t *a, size_t n, uint32_t *bingcd, uint32_t *fermat) {
    for (size_t i = 0; i < n; i++) {
        fp x, y, z;
        for (int j = 0; j < 8; j++) x.v[j] = a[8 * i + j];
        fp_inv_bingcd(y, x);
        fp_inv_fermat(z, x);
        for (int j = 0; j < 8; j++) {
            bingcd[8 * i + j] = y.v[j];
            fermat[8 * i + j] = z.v[j];
        }
    }
}
// subgroup membership as the kernels evaluate it (G1: on-curve; G2: the psi criterion), from wire bytes
EXPORT void sim_torsion_free(const uint8_t *g1pts, size_t n1, const uint8_t *g2pts, size_t n2, uint8_t *out1, uint8_t *out2) {
    for (size_t i = 0; i < n1; i++) {
        g1m p;
        uint8_t fl;
        item_g1_decode(p, fl, g1pts + 64 * i, true);
        out1[i] = item_g1_torsion_free(p, fl);
    }
    for (size_t i = 0; i < n2; i++) {
        g2m q;
        uint8_t fl;
        item_g2_decode(q, fl, g2pts + 128 * i, true);
        out2[i] = item_g2_torsion_free(q, fl);
    }
}
// randomised batch verification: the per-tuple scalar and the two scalings of items.cuh
EXPORT void sim_rbv_scalar(const uint8_t *seed, uint64_t idx, int nbits, uint32_t *kw) {
    rbv_k kk;
    rbv_scalar(kk, seed, idx, nbits);
    kw[0] = kk.a[0]; kw[1] = kk.a[1]; kw[2] = kk.b[0]; kw[3] = kk.b[1];
}
EXPORT void sim_rbv_scale(const uint8_t *g1pts, const uint8_t *g2pts, size_t n, const uint8_t *seed, uint64_t idx0, int nbits, uint8_t *out_g1_affine,
                          uint8_t *out_g1_sum, uint8_t *out_g2_sum) {
    g1p s1;
    g2p s2;
    pt_set_identity(s1);
    pt_set_identity(s2);
    for (size_t i = 0; i < n; i++) {
        rbv_k kw;
        rbv_scalar(kw, seed, idx0 + i, nbits);
        g1m p;
        g2m q;
        uint8_t fl, flq;
        item_g1_decode(p, fl, g1pts + 64 * i, false);
        item_g2_decode(q, flq, g2pts + 128 * i, false);
        g1p t1;
        g2p t2;
        item_rbv_scale_proj<F1, g1m>(t1, p, fl, kw, nbits);
        item_rbv_scale_proj<F2, g2m>(t2, q, flq, kw, nbits);
        pt_add(s1, s1, t1);
        pt_add(s2, s2, t2);
        item_rbv_scale_g1(p, fl, kw, nbits);
        item_g1_encode(out_g1_affine + 64 * i, p, fl);
    }
    g1m a1;
    g2m a2;
    bool inf;
    pt_to_affine(a1.x, a1.y, inf, s1);
    item_g1_encode(out_g1_sum, a1, inf ? FL_INF : 0);
    pt_to_affine(a2.x, a2.y, inf, s2);
    item_g2_encode(out_g2_sum, a2, inf ? FL_INF : 0);
}
EXPORT void sim_g1_codec(const uint8_t *pts, size_t n, uint8_t *comp, uint8_t *back, uint8_t *status) {
    for (size_t i = 0; i < n; i++) {
        g1m p, r;
        uint8_t fl, fl2;
        item_g1_decode(p, fl, pts + 64 * i, false);
        item_g1_compress(comp + 32 * i, p, fl);
        item_g1_decompress(r, fl2, comp + 32 * i);
        item_g1_encode(back + 64 * i, r, fl2);
        status[i] = !(fl2 & FL_BAD);
    }
}
EXPORT void sim_g2_codec(const uint8_t *pts, size_t n, uint8_t *comp, uint8_t *back, uint8_t *status) {
    for (size_t i = 0; i < n; i++) {
        g2m p, r;
        uint8_t fl, fl2;
        item_g2_decode(p, fl, pts + 128 * i, false);
        item_g2_compress(comp + 64 * i, p, fl);
        item_g2_decompress(r, fl2, comp + 64 * i);
        item_g2_encode(back + 128 * i, r, fl2);
        status[i] = !(fl2 & FL_BAD);
    }
}
// wide product / reduction / Fp2 product in isolation (raw limbs)
EXPORT void sim_fp2_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    fp2 x, y, z;
    memcpy(&x, a, 64);
    memcpy(&y, b, 64);
    fp2_mul(z, x, y);
    memcpy(r, &z, 64);
}
EXPORT void sim_fp_dot2(uint32_t *r, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d) {
    fp x, y, z, w, o;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    memcpy(z.v, c, 32);
    memcpy(w.v, d, 32);
    fp_dot2_inl(o, x, y, z, w);
    memcpy(r, o.v, 32);
}
// ---- lane emulation: host threads run the same body; the two lanes of a pair exchange through a rendezvous, and (wide slot
//      machine: 32 lanes = 16 pairs) the whole emulated warp can synchronise ----
#include <pthread.h>
static thread_local int tl_lane = 0;
static pthread_barrier_t g_bar[16], g_wbar;
static uint32_t g_slot[32][16];
extern "C" uint32_t bn_host_lane_par() { return (uint32_t)(tl_lane & 1); }
extern "C" int bn_host_lane_id() { return tl_lane; }
extern "C" void bn_host_warp_sync() { pthread_barrier_wait(&g_wbar); }
extern "C" void bn_host_lane_xchg(uint32_t *out, const uint32_t *in, int nwords) {
    memcpy(g_slot[tl_lane], in, 4 * nwords);
    pthread_barrier_wait(&g_bar[tl_lane >> 1]);
    memcpy(out, g_slot[tl_lane ^ 1], 4 * nwords);
    pthread_barrier_wait(&g_bar[tl_lane >> 1]);
}
struct pair_job { void (*fn)(void *); void *arg; int lane; };
static void *pair_main(void *p) {
    pair_job *j = (pair_job *)p;
    tl_lane = j->lane;
    j->fn(j->arg);
    return nullptr;
}
static void run_lanes(void (*fn)(void *), void *arg, int nlanes) {
    for (int i = 0; i < nlanes / 2; i++) pthread_barrier_init(&g_bar[i], nullptr, 2);
    pthread_barrier_init(&g_wbar, nullptr, nlanes);
    pthread_t th[32];
    pair_job jobs[32];
    for (int i = 0; i < nlanes; i++) {
        jobs[i] = {fn, arg, i};
        pthread_create(&th[i], nullptr, pair_main, &jobs[i]);
    }
    for (int i = 0; i < nlanes; i++) pthread_join(th[i], nullptr);
    for (int i = 0; i < nlanes / 2; i++) pthread_barrier_destroy(&g_bar[i]);
    pthread_barrier_destroy(&g_wbar);
}
static void run_pair(void (*fn)(void *), void *arg) { run_lanes(fn, arg, 2); }
// lane-pair versions of the pairing bodies (both emulated lanes run the item function)
struct l2_args { const uint8_t *g1, *g2; const uint8_t *p0, *q0, *p1, *q1; int fixed; int mode; fp12 F; uint8_t fl; };
static void l2_pairing_body(void *vp) {
    l2_args *a = (l2_args *)vp;
    g1m p;
    g2m q;
    uint8_t flp, flq, fl;
    item_g1_decode(p, flp, a->g1, false);
    item_g2_decode(q, flq, a->g2, false);
    item_miller1_l2(a->F, fl, p, flp, q, flq);
    if (a->mode == 0) {
        uint32_t dummy_in = 0, dummy_out;
        bn_host_lane_xchg(&dummy_out, &dummy_in, 1);  // rendezvous: both halves of F are stored before either lane reloads it
        item_final_exp_l2(a->F);
    }
    if (bn_host_lane_par() == 0) a->fl = fl;
}
EXPORT void sim_pairing_batch_l2(const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *out, uint8_t *status, int mode) {
    for (size_t i = 0; i < n; i++) {
        l2_args a;
        memset(&a, 0, sizeof a);
        a.g1 = g1 + 64 * i;
        a.g2 = g2 + 128 * i;
        a.mode = mode;
        run_pair(l2_pairing_body, &a);
        if (a.fl & FL_BAD) memset(out + 384 * i, 0, 384);
        else fp12_to_bytes(out + 384 * i, a.F);
        if (status) status[i] = !(a.fl & FL_BAD);
    }
}
static void l2_miller2_body(void *vp) {
    l2_args *x = (l2_args *)vp;
    g1m a, c;
    g2m b, d;
    uint8_t fa, fb = 0, fc, fd, fl;
    item_g1_decode(a, fa, x->p0, false);
    if (!x->fixed) item_g2_decode(b, fb, x->q0, false);
    item_g1_decode(c, fc, x->p1, false);
    item_g2_decode(d, fd, x->q1, false);
    if (x->fixed) item_miller2_l2<true>(x->F, fl, a, fa, nullptr, 0, c, fc, d, fd);
    else item_miller2_l2<false>(x->F, fl, a, fa, &b, fb, c, fc, d, fd);
}
EXPORT void sim_miller2_l2(const uint8_t *p0, const uint8_t *q0, const uint8_t *p1, const uint8_t *q1, int fixed, uint8_t *out) {
    l2_args a;
    memset(&a, 0, sizeof a);
    a.p0 = p0; a.q0 = q0; a.p1 = p1; a.q1 = q1; a.fixed = fixed;
    run_pair(l2_miller2_body, &a);
    fp12_to_bytes(out, a.F);
}
static void l2_fe_body(void *vp) { item_final_exp_l2(*(fp12 *)vp); }
EXPORT void sim_final_exp_l2(const uint8_t *in, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) {
        fp12 f;
        fp12_from_bytes(f, in + 384 * i);
        run_pair(l2_fe_body, &f);
        fp12_to_bytes(out + 384 * i, f);
    }
}

// the same Fp12 operations as sim_fp12_op in the LANE-PAIR representation (two emulated lanes, tower.cuh: fp2l)
struct l2_op_args { int op; fp12 x, y, r; };
static void l2_op_body(void *vp) {
    l2_op_args *a = (l2_op_args *)vp;
    fp12l x, y, r;
    fp12_load(x, a->x);
    fp12_load(y, a->y);
    switch (a->op) {
        case 0: fp12_mul(r, x, y); break;
        case 1: fp12_sqr(r, x); break;
        case 2: fp12_inv(r, x); break;
        case 3: fp12_conj(r, x); break;
        case 4: case 5: case 6: fp12_frob(r, x, a->op - 3); break;
        case 7: cyclotomic_square(r, x); break;
        case 9: r = x; fp12_mul_by_034_034(r, y.c0.c0, y.c0.c1, y.c0.c2, y.c1.c0, y.c1.c1, y.c1.c2); break;
        default: r = x; fp12_mul_by_034(r, y.c0.c0, y.c0.c1, y.c0.c2); break;
    }
    fp12_store(a->r, r);
}
EXPORT int sim_fp12_op_l2(int op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    l2_op_args args;
    args.op = op;
    if (!fp12_from_bytes(args.x, a)) return -1;
    if (b) {
        if (!fp12_from_bytes(args.y, b)) return -1;
    } else {
        args.y = args.x;
    }
    run_pair(l2_op_body, &args);
    fp12_to_bytes(out, args.r);
    return 0;
}


// Executed multiplier work (INT32 multiply-add operations, fp.cuh: BN_COUNT_OPS) of ONE kernel body on ONE item, wire
// codecs excluded.  kind: 0 hash_to_g1, 1 hash_to_g2 (count 2), 2 one-pair Miller loop, 3 two-pair Miller loop with the
// fixed -g2 table, 4 two-pair Miller loop (both variable), 5 final exponentiation (of the kind-2 result), 6 G2 subgroup check.
EXPORT unsigned long long sim_count_item(int kind, const uint8_t *p0, const uint8_t *q0, const uint8_t *p1, const uint8_t *q1, const uint8_t *msg,
                                         size_t mlen, const uint8_t *dst, size_t dst_len) {
    g1m a, c;
    g2m b, d;
    uint8_t fa = 0, fb = 0, fc = 0, fd = 0, fl;
    fp12 f;
    if (kind >= 2) {
        item_g1_decode(a, fa, p0, false);
        item_g2_decode(b, fb, q0, false);
        item_g1_decode(c, fc, p1, false);
        item_g2_decode(d, fd, q1, false);
    }
    if (kind == 5) item_miller1(f, fl, a, fa, b, fb);
    g_ops.store(0);
    switch (kind) {
        case 0: item_hash_g1(a, fl, msg, mlen, dst, (uint32_t)dst_len, 2); break;
        case 1: item_hash_g2(b, fl, msg, mlen, dst, (uint32_t)dst_len, 2); break;
        case 2: item_miller1(f, fl, a, fa, b, fb); break;
        case 3: item_miller2<true>(f, fl, a, fa, nullptr, 0, c, fc, d, fd); break;
        case 4: item_miller2<false>(f, fl, a, fa, &b, fb, c, fc, d, fd); break;
        case 5: item_final_exp(f); break;
        case 6: (void)item_g2_torsion_free(b, fb); break;
        default: break;
    }
    return g_ops.load();
}

// ---- the slot machine (vm.cuh) on two emulated lanes: fam 0 one pair (p1, q1), 1 two pairs with the fixed -g2 table
//      (p0; p1, q1), 2 two variable pairs; mode bit 0: also run the final exponentiation.  Wire bytes in, 384 wire bytes out.
#include "../../bls_bn254_b200/csrc/vm.cuh"
struct vm_args { int fam, fexp; g1m p0, p1; g2m q0, q1; fp12 F; fp2 scratch[vm::FEXP_SCRATCH_FP2]; };
static void vm_body(void *vp) {
    vm_args *a = (vm_args *)vp;
    fp slots[64];
    const vm::Slots M = {slots};
    vm::Env E = {};
    E.p0 = &a->p0.x; E.p1 = &a->p1.x;
    E.q0 = &a->q0.x; E.q1 = &a->q1.x;
    E.scratch = a->scratch;
    E.scratch_stride = 1;
    const bool odd = bn_host_lane_par() != 0;
    if (a->fam >= 0) vm::miller(M, E, a->fam);
    else vm::load_f12(M, a->F, odd);
    if (a->fexp) vm::final_exp(M, E);
    uint32_t dummy_in = 0, dummy_out;
    bn_host_lane_xchg(&dummy_out, &dummy_in, 1);
    vm::store_f12(a->F, M, odd);
}
EXPORT void sim_vm(int fam, int fexp, const uint8_t *p0, const uint8_t *q0, const uint8_t *p1, const uint8_t *q1, const uint8_t *f_in, uint8_t *out) {
    static vm_args a;
    memset(&a, 0, sizeof a);
    uint8_t fl;
    a.fam = fam;
    a.fexp = fexp;
    if (p0) item_g1_decode(a.p0, fl, p0, 0);
    if (q0) item_g2_decode(a.q0, fl, q0, 0);
    if (p1) item_g1_decode(a.p1, fl, p1, 0);
    if (q1) item_g2_decode(a.q1, fl, q1, 0);
    if (f_in) fp12_from_bytes(a.F, f_in);
    run_pair(vm_body, &a);
    fp12_to_bytes(out, a.F);
}
EXPORT unsigned long long sim_vm_ops(int fam, int fexp, const uint8_t *p0, const uint8_t *q0, const uint8_t *p1, const uint8_t *q1, const uint8_t *f_in) {
    uint8_t out[384];
    g_ops.store(0);
    sim_vm(fam, fexp, p0, q0, p1, q1, f_in, out);
    return g_ops.load();
}

// ---- the WIDE slot machine (vm.cuh: run on WSlots) on 32 emulated lanes: same arguments and meaning as sim_vm ----
struct vmw_args { int fam, fexp; g1m p0, p1; g2m q0, q1; fp12 F; fp (*slots)[2]; };
static void vmw_body(void *vp) {
    vmw_args *a = (vmw_args *)vp;
    const vm::WSlots M = {a->slots};
    vm::Env E = {};
    if (a->fam >= 0) {
        vm::wide_set_item(M, a->fam == 0 ? nullptr : &a->p0.x, a->fam == 2 ? &a->q0.x : nullptr, &a->p1.x, &a->q1.x);
        vm::miller(M, E, a->fam);
    } else {
        vm::load_f12(M, a->F);
    }
    if (a->fexp) vm::final_exp(M, E);
    vm::wide_sync();
    vm::store_f12(a->F, M);
}
EXPORT void sim_vm_wide(int fam, int fexp, const uint8_t *p0, const uint8_t *q0, const uint8_t *p1, const uint8_t *q1, const uint8_t *f_in, uint8_t *out) {
    static vmw_args a;
    static fp slots[vm::W_NSLOT][2];
    memset(&a, 0, sizeof a);
    memset(slots, 0, sizeof slots);
    uint8_t fl;
    a.fam = fam;
    a.fexp = fexp;
    a.slots = slots;
    if (p0) item_g1_decode(a.p0, fl, p0, 0);
    if (q0) item_g2_decode(a.q0, fl, q0, 0);
    if (p1) item_g1_decode(a.p1, fl, p1, 0);
    if (q1) item_g2_decode(a.q1, fl, q1, 0);
    if (f_in) fp12_from_bytes(a.F, f_in);
    run_lanes(vmw_body, &a, 32);
    fp12_to_bytes(out, a.F);
}

// the wide machine's lazy linear combination: sum_i coef[i] * val[i] mod p through the unreduced accumulator (plain integers in
// and out, 32 bytes big-endian each; |coef| <= 40, n <= 12 -- the bound the interpreter's gathers stay within)
EXPORT void sim_wide_lincomb(const uint8_t *vals, const int *coefs, int n, uint8_t *out) {
    vm::WAcc A;
    vm::wacc_zero(A);
    for (int i = 0; i < n; i++) {
        fp v;
        fp_raw_from_be(v, vals + 32 * i);
        vm::wacc_mad(A, v, (uint32_t)(coefs[i] < 0 ? -coefs[i] : coefs[i]), coefs[i] < 0);
    }
    fp r = vm::wacc_reduce(A);
    for (int i = 0; i < 8; i++) {
        uint8_t *q = out + 4 * (7 - i);
        q[0] = (uint8_t)(r.v[i] >> 24); q[1] = (uint8_t)(r.v[i] >> 16); q[2] = (uint8_t)(r.v[i] >> 8); q[3] = (uint8_t)r.v[i];
    }
}

// ---- G2Prepared: the line tables (wire form as the oracle's orc_g2_prepare) and the two-pair loop fed from them ----------
EXPORT void sim_g2_prepare(const uint8_t *pt, uint8_t *out) {
    static line_t<fp2> tab[BN254_NUM_LINES];
    g2m q;
    uint8_t fl;
    item_g2_decode(q, fl, pt, 0);
    item_g2_prepare(tab, q, fl);
    const fp2 *c = reinterpret_cast<const fp2 *>(tab);
    for (int i = 0; i < BN254_NUM_LINES * 3; i++) {
        fp_to_be(out + 64 * i, c[i].c0);
        fp_to_be(out + 64 * i + 32, c[i].c1);
    }
}
struct prep_args { g1m p0, p1; uint8_t f0, f1, fq; line_t<fp2> tab[BN254_NUM_LINES]; fp12 F; uint8_t fl; };
static void prep_l2_body(void *vp) {
    prep_args *a = (prep_args *)vp;
    uint8_t fl;
    item_miller2_prep_l2(a->F, fl, a->p0, a->f0, a->p1, a->f1, a->tab, a->fq);
    if (bn_host_lane_par() == 0) a->fl = fl;
}
// lanes = 0: one thread per pairing, 1: two emulated lanes
EXPORT void sim_miller2_prepared(const uint8_t *p0, const uint8_t *p1, const uint8_t *q1, int lanes, uint8_t *out) {
    static prep_args a;
    memset(&a, 0, sizeof a);
    g2m q;
    item_g1_decode(a.p0, a.f0, p0, 0);
    item_g1_decode(a.p1, a.f1, p1, 0);
    item_g2_decode(q, a.fq, q1, 0);
    item_g2_prepare(a.tab, q, a.fq);
    if (lanes) run_pair(prep_l2_body, &a);
    else item_miller2_prep(a.F, a.fl, a.p0, a.f0, a.p1, a.f1, a.tab, a.fq);
    fp12_to_bytes(out, a.F);
}
