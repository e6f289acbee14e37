/* bls_bn254_b200.h -- C ABI of libbls_bn254_b200.so: batch BLS verification over BN254 on B200.
 *
 * The reference (mikelodder7/bls-bn254) has no FFI; its boundary is the Rust API of
 * src/inner_types.  Every entry point below is the BATCH form of one reference function and uses
 * the reference's own byte formats, so a Rust shim maps 1:1:
 *
 *   G1 point   64 B  x || y, 32-byte big-endian each            g1.rs:296-302 (to_uncompressed)
 *   G2 point  128 B  x.c1 || x.c0 || y.c1 || y.c0               g2.rs:290-300
 *   G1 / G2 compressed 32 B / 64 B, flag in bit 7 of byte 0     g1.rs:283-288, g2.rs:274-283
 *   Gt / MillerLoopResult 384 B  c0.c0.c0 || c0.c0.c1 || ...    pairings.rs:499-514 (Gt::to_repr)
 *   identity = (0, 1) uncompressed                              g1.rs:265-271
 *   messages: one byte blob + u64 offsets[n+1] (message i = msgs[offsets[i] .. offsets[i+1]))
 *
 * Return value: 0 on success; -1..-4 = Bn254Error::{InvalidScalarBytes, InvalidG1Bytes,
 * InvalidG2Bytes, InvalidGtBytes} (error.rs:5-10) for malformed call-level input; -10 bad argument;
 * -20 no usable CUDA device (there is NO CPU fallback); -100-k = CUDA error k (k < 1000); -2000-k = NCCL error k
 * (-2000 itself: libnccl could not be loaded / no communicator; SURVEY.md 8(b) proposed -200-k, which collides with CUDA error
 * numbers above 100, e.g. 700 "illegal address").  Per-item problems
 * (undecodable point, SvdW exception where the reference panics: fp.rs:325, fp2.rs:255) never abort
 * a batch: the item's status byte is 0 and its output is all zero bytes (ok[i] = 0 for verify).
 *
 * Threading: calls are synchronous; a ctx is single-caller; several contexts may coexist.
 * Ownership: the caller owns every buffer; nothing is retained after return.
 * Host-buffer entry points shard items across the ctx's devices by contiguous index range (no
 * inter-GPU traffic) and stream each device's share through a two-stage chunk pipeline (upload of
 * chunk k+1 and download of chunk k-1 overlap the kernels of chunk k).  Pinned caller buffers
 * (cudaHostAlloc / cudaHostRegister / bn254_host_register) make every copy asynchronous; pageable
 * buffers work and are staged.  "_dev" entry points take device pointers resident on the ctx's FIRST device.
 *
 * Latency note: a lone verification through this library takes ~4 ms (a single pairing is a latency-
 * bound chain of ~17 000 dependent field multiplications; launches of at most "wide_max" = 1536 pairings
 * give each pairing a whole warp, which cuts the chain 2-3 x; a CPU core does it in ~2 ms).  The library
 * is for batches: throughput saturates from ~2^14 tuples per call.
 *
 * Identity policy of the verification predicates: a public key or (per-tuple) signature that decodes
 * as the identity -- x == 0, or y == 1 under the reference's decode convention (g1.rs:318-324) -- is
 * rejected like an undecodable one (KeyValidate of the IETF BLS draft); an aggregate signature may be
 * the identity.  The pairing / Miller-loop entry points keep the reference's semantics (an identity
 * input contributes the factor 1, pairings.rs:789-800).
 */
#ifndef BLS_BN254_B200_H
#define BLS_BN254_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bn254_ctx bn254_ctx;

#define BN254_OK 0
#define BN254_ERR_INVALID_SCALAR_BYTES (-1)
#define BN254_ERR_INVALID_G1_BYTES (-2)
#define BN254_ERR_INVALID_G2_BYTES (-3)
#define BN254_ERR_INVALID_GT_BYTES (-4)
#define BN254_ERR_ARG (-10)
#define BN254_ERR_NO_DEVICE (-20)
#define BN254_ERR_CUDA_BASE (-100)
#define BN254_ERR_NCCL_BASE (-2000)

#define BN254_MIN_SIG 0 /* pk in G2 (128 B), sig in G1 (64 B), messages hashed to G1 */
#define BN254_MIN_PK 1  /* pk in G1 (64 B), sig in G2 (128 B), messages hashed to G2 */
#define BN254_FLAG_CHECK_ON_CURVE 1u /* reject keys / signatures that are not on the curve */
#define BN254_FLAG_CHECK_SUBGROUP 2u /* also require keys / signatures in the order-r subgroup (implies on-curve); every verify / aggregate entry point */

/* ---- context ---------------------------------------------------------------------------------- */
int bn254_device_count(void);
/* device_ids == NULL: devices 0..n_dev-1.  Fails with BN254_ERR_NO_DEVICE when CUDA is unavailable. */
int bn254_ctx_create(bn254_ctx **out, const int *device_ids, int n_dev);
void bn254_ctx_destroy(bn254_ctx *ctx);
const char *bn254_last_error(const bn254_ctx *ctx);
/* Tuning / test knobs of a live ctx (the BN254_LANES, BN254_WIDE_MAX, BN254_CHUNK, BN254_SPLIT environment variables set the
 * same values at creation): "lanes" = shape of the pairing kernels -- 0 picked per launch (default), 1 one thread per pairing,
 * 2 two lanes per pairing, 3 the slot machine (two lanes, operands in shared memory), 4 the wide slot machine (one WARP per
 * pairing: the latency shape); results are bit-identical in every shape (the parity suite runs all of them); "wide_max" =
 * with "lanes" 0, launches of at most this many pairings use the wide shape; "chunk" = items staged per host<->device round
 * (default 2^18); "split" = concurrent slices per verification batch (default 1); "sk_msm_min" = average tuples per key from
 * which the same-key path sums by bucket-method multi-scalar multiplications instead of scaling every point (default 128,
 * 0 = never; same verdicts either way). */
int bn254_ctx_set_option(bn254_ctx *ctx, const char *name, long value);
/* Pin / unpin caller memory (cudaHostRegister) so that the chunk pipeline copies it asynchronously. */
int bn254_host_register(void *p, size_t bytes);
int bn254_host_unregister(void *p);
/* ---- NCCL communicator for jobs of one process per GPU (the aggregate path's single exchange).
 * Rank 0 calls bn254_comm_unique_id and ships the 128 bytes to the other ranks by any means (torch.distributed, MPI, a
 * socket); every rank then calls bn254_comm_init_rank on its one-device ctx.  A ctx created with several devices owns one
 * communicator per device already (ncclCommInitAll).  libnccl.so.2 is bound at run time. */
#define BN254_COMM_ID_BYTES 128
int bn254_comm_unique_id(uint8_t *out /* 128 */);
int bn254_comm_init_rank(bn254_ctx *ctx, const uint8_t *id /* 128 */, int nranks, int rank);
int bn254_comm_nranks(const bn254_ctx *ctx);
/* number of kernels this ctx has launched so far (all devices; an NCCL collective counts as one) */
uint64_t bn254_launch_count(const bn254_ctx *ctx);
/* cudaStream_t of the ctx's first device (for event timing of the _dev entry points) */
void *bn254_stream(bn254_ctx *ctx);
int bn254_sync(bn254_ctx *ctx);
/* Per-launch timing: when enabled every kernel launch is bracketed by a CUDA event pair on the
 * launch stream; bn254_profile_drain synchronises and writes one line per launch, device after
 * device, "kernel_name items milliseconds\n", into buf (BN254_ERR_ARG if it does not fit). */
int bn254_profile_enable(bn254_ctx *ctx, int on);
int bn254_profile_drain(bn254_ctx *ctx, char *buf, size_t cap);

/* ---- hash to curve: G1Projective::hash / ::encode (g1.rs:910-928), G2Projective::hash / ::encode
 *      (g2.rs:919-936) + to_uncompressed.  status may be NULL. ------------------------------------ */
int bn254_hash_to_g1(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                     uint8_t *out /* n*64 */, uint8_t *status /* n */);
int bn254_encode_to_g1(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                       uint8_t *out, uint8_t *status);
int bn254_hash_to_g2(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                     uint8_t *out /* n*128 */, uint8_t *status);
int bn254_encode_to_g2(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                       uint8_t *out, uint8_t *status);

/* ---- pairing engine: pairing() (pairings.rs:760-802), multi_miller_loop (808-857),
 *      MillerLoopResult::final_exponentiation (50-178), MillerLoopResult + (37-43) --------------- */
int bn254_pairing_batch(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *gt /* n*384 */, uint8_t *status);
int bn254_miller_loop_batch(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *ml /* n*384 */, uint8_t *status);
int bn254_multi_miller_loop(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *ml /* 384 */);
int bn254_final_exponentiation(bn254_ctx *ctx, const uint8_t *ml, size_t n, uint8_t *gt, uint8_t *status);
int bn254_fp12_product(bn254_ctx *ctx, const uint8_t *in /* n*384 */, size_t n, uint8_t *out /* 384 */);

/* ---- BLS predicates (absent from the reference snapshot; defined in SURVEY.md 8(a) row V) ---------
 * min-sig: e(sig, -g2) * e(H_G1(m), pk) == 1;  min-pk: e(-g1, sig) * e(pk, H_G2(m)) == 1
 * aggregate (distinct messages): e(aggsig, -g2) * prod e(H(m_i), pk_i) == 1  (resp. min-pk) */
int bn254_verify_batch(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, const uint8_t *sigs,
                       size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok /* n bytes 0/1 */);
/* Randomised batch verification (no counterpart in the reference).  The tuples are
 * taken in consecutive groups of `group` (1..BN254_RBV_MAX_GROUP; the last group may be short) and every group
 * gets ONE accept byte from a small-exponent test:
 *     min-sig: e(sum_i r_i sig_i, -g2) * prod_i e(r_i H_G1(m_i), pk_i) == 1       (min-pk symmetric)
 * i.e. one Miller loop per tuple and one final exponentiation per GROUP instead of per tuple.
 * r_i = a_i + b_i * lambda mod r with (a_i, b_i) = the first / second rand_bits/2 bits (rand_bits = 64 or 128) of
 * SHA-256(seed[32] || LE64(i)), big-endian, (0, 0) -> (1, 0); i = index of the tuple in this call and
 * lambda = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd, the eigenvalue of (x, y) -> (beta x, y) on the
 * order-r subgroups.  r_i takes 2^rand_bits distinct values ((a, b) -> a + b lambda is injective below 2^64) and
 * [r_i]P costs only rand_bits/2 doublings.  A group whose tuples all verify is always accepted; a group with an invalid tuple is
 * accepted with probability <= 2^-rand_bits PROVIDED the seed is unpredictable to whoever produced the
 * signatures (fresh CSPRNG output per call) and public keys / signatures lie in the order-r subgroups
 * (BN254_FLAG_CHECK_SUBGROUP for untrusted min-sig public keys / min-pk signatures).  On a rejected group the
 * caller falls back to bn254_verify_batch on that group's tuples for per-tuple verdicts. */
#define BN254_RBV_MAX_GROUP 65536
int bn254_verify_batch_randomized(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets,
                                  const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags,
                                  const uint8_t *seed /* 32 B */, int rand_bits, size_t group,
                                  uint8_t *ok_groups /* ceil(n/group) bytes 0/1 */);
/* Same-key randomised verification ("same-key grouping"; no counterpart in the reference): many messages
 * signed under FEW keys (a validator that signs a stream of messages).  Key k owns the tuples [key_offsets[k], key_offsets[k+1])
 * of `msgs` / `sigs` (the caller sorts its tuples by key) and gets ONE accept byte:
 *     min-sig: e(sum_i r_i sig_i, -g2) * e(sum_i r_i H_G1(m_i), pk_k) == 1     min-pk: e(-g1, sum_i r_i sig_i) * e(pk_k, sum_i r_i H_G2(m_i)) == 1
 * (bilinearity: prod_i e(r_i H(m_i), pk_k) = e(sum_i r_i H(m_i), pk_k)).  Per tuple that is one hash to the curve and two short
 * scalar multiplications; the pairing work -- one two-pair Miller loop and one final exponentiation -- is per KEY.  r_i is the
 * scalar of bn254_verify_batch_randomized for tuple index i, the error bound and its preconditions (unpredictable seed, inputs in
 * the order-r subgroups) are the same.  A key without tuples, an unusable key (undecodable / identity / failing the checks of
 * `flags`) or an unusable signature rejects that key.  On a rejected key the caller falls back to bn254_verify_batch on the key's
 * tuples.  A key is never split over chunks: device scratch grows to ~0.5 KB per tuple of the largest key.
 * `_dev`: device pointers on device 0 for everything except key_offsets (host; key_offsets[0] must be 0). */
int bn254_verify_same_key_randomized(bn254_ctx *ctx, int variant, const uint8_t *pks /* n_keys records */,
                                     const uint64_t *key_offsets /* n_keys+1, in tuples */, size_t n_keys, const uint8_t *msgs,
                                     const uint64_t *offsets /* key_offsets[n_keys]+1 */, const uint8_t *sigs, const uint8_t *dst,
                                     size_t dst_len, uint32_t flags, const uint8_t *seed /* 32 B */, int rand_bits,
                                     uint8_t *ok_keys /* n_keys bytes 0/1 */);
int bn254_verify_same_key_randomized_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint64_t *key_offsets /* HOST */,
                                         size_t n_keys, const uint8_t *msgs, const uint64_t *offsets, const uint8_t *sigs,
                                         const uint8_t *dst, size_t dst_len, uint32_t flags, const uint8_t *seed, int rand_bits,
                                         uint8_t *ok_keys);
/* FastAggregateVerify of the IETF BLS draft (many signers, ONE message per group), n_groups groups at once: group g = the public
 * keys [key_offsets[g], key_offsets[g+1]) of `pks`, message g, aggregate signature g:
 *     min-sig: e(aggsig_g, -g2) * e(H_G1(m_g), sum pk) == 1          min-pk: e(-g1, aggsig_g) * e(sum pk, H_G2(m_g)) == 1
 * The aggregate key of every group is summed on the GPU; then one ordinary verification per GROUP.  A group with no key, with an
 * undecodable / identity / (per flags) invalid key, or whose keys sum to the identity is rejected. */
int bn254_fast_aggregate_verify_batch(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint64_t *key_offsets /* n_groups+1 */,
                                      const uint8_t *msgs, const uint64_t *offsets, const uint8_t *agg_sigs, size_t n_groups,
                                      const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok /* n_groups bytes 0/1 */);
/* ---- prepared public keys: G2Prepared (pairings.rs:609-660, 725-757) for min-sig keys.  bn254_g2_prepare computes the 88 line
 * triples of every key once (16 896 bytes per key, resident on every device of the ctx; `flags` = the validation applied to the
 * keys: on-curve / subgroup; identity and undecodable keys are remembered as unusable).  bn254_verify_batch_prepared verifies
 * tuple i against key key_index[i]: the Miller loop then takes BOTH line sequences from tables (the fixed -g2 and the key's) and
 * does no twist-point arithmetic -- ~20 % less work per verification for a validator set that signs many messages; `flags`
 * there applies to the signatures.  Accept bytes are identical to bn254_verify_batch on the expanded key array.
 * bn254_g2_prepared_export returns the tables as 88 x 3 Fp2 values per key (c0 || c1, 32-byte big-endian each) for inspection
 * / caching; status 0 and zero bytes for an unusable key. */
typedef struct bn254_g2_prepared bn254_g2_prepared;
int bn254_g2_prepare(bn254_ctx *ctx, const uint8_t *pks /* n*128 */, size_t n, uint32_t flags, bn254_g2_prepared **out);
void bn254_g2_prepared_free(bn254_g2_prepared *pr);
size_t bn254_g2_prepared_count(const bn254_g2_prepared *pr);
int bn254_g2_prepared_export(bn254_ctx *ctx, const bn254_g2_prepared *pr, size_t first, size_t count, uint8_t *out /* count*88*192 */,
                             uint8_t *status /* count, may be NULL */);
int bn254_verify_batch_prepared(bn254_ctx *ctx, const bn254_g2_prepared *pr, const uint32_t *key_index /* n */, const uint8_t *msgs,
                                const uint64_t *offsets, const uint8_t *sigs /* n*64 */, size_t n, const uint8_t *dst, size_t dst_len,
                                uint32_t flags, uint8_t *ok /* n bytes 0/1 */);
int bn254_aggregate_verify(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                           const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok /* 1 byte */);
/* One shard of an aggregate verification: the Fp12 product of this shard's Miller loops (384 B,
 * un-exponentiated) and a "bad input" byte.  agg_sig is given to exactly one shard (NULL elsewhere).
 * The shards' partials are exchanged by the caller (one small all-gather) and finished with
 * bn254_aggregate_finish: product, ONE final exponentiation, compare with 1. */
int bn254_aggregate_partial(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                            const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *partial /* 384 */,
                            uint8_t *bad /* 1 */);
int bn254_aggregate_finish(bn254_ctx *ctx, const uint8_t *partials /* k*384 */, const uint8_t *bad /* k */, size_t k, uint8_t *ok);
/* The whole predicate for a job of one process per GPU: every rank passes ITS shard of (pk_i, m_i); agg_sig is read on rank 0
 * only.  Per rank: hash -> Miller loops -> product tree whose last pass writes the NCCL send record; ONE ncclAllGather of 400
 * bytes per rank over NVLink on the library stream; ONE kernel (<= nranks products, the single final exponentiation, compare).
 * *ok = the job's verdict, on every rank.  Needs bn254_comm_init_rank when nranks > 1. */
int bn254_aggregate_verify_dist(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok /* 1 byte */);

/* ---- codecs and predicates: to_/from_compressed (g1.rs:283-328, g2.rs:274-340), is_on_curve ---- */
int bn254_g1_compress(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* n*32 */);
int bn254_g1_decompress(bn254_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out /* n*64 */, uint8_t *status);
int bn254_g2_compress(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* n*64 */);
int bn254_g2_decompress(bn254_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out /* n*128 */, uint8_t *status);
int bn254_g1_is_on_curve(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* n */);
int bn254_g2_is_on_curve(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* n */);
/* is_torsion_free (g1.rs:852-855, g2.rs:733-736): 1 for a point on the curve with [r]P = identity */
int bn254_g1_is_torsion_free(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* n */);
int bn254_g2_is_torsion_free(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* n */);

/* ---- scalar multiplication and sums (sign / keygen / aggregate side; g1.rs:821-841,
 *      g2.rs:866-886).  scalars: n x 32-byte big-endian integers. -------------------------------- */
int bn254_g1_mul(bn254_ctx *ctx, const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out);
int bn254_g2_mul(bn254_ctx *ctx, const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out);
int bn254_g1_sum(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* 64 */);
int bn254_g2_sum(bn254_ctx *ctx, const uint8_t *pts, size_t n, uint8_t *out /* 128 */);

/* ---- device-resident variants (all pointers are device pointers on the ctx's first device; work is
 *      enqueued on bn254_stream(ctx) and NOT synchronised -- call bn254_sync or use events) -------- */
int bn254_hash_to_g1_dev(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                         int count, uint8_t *out, uint8_t *status);
int bn254_hash_to_g2_dev(bn254_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                         int count, uint8_t *out, uint8_t *status);
int bn254_pairing_batch_dev(bn254_ctx *ctx, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *gt, uint8_t *status);
int bn254_verify_batch_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets,
                           const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok);
int bn254_verify_batch_prepared_dev(bn254_ctx *ctx, const bn254_g2_prepared *pr, const uint32_t *key_index, const uint8_t *msgs,
                                    const uint64_t *offsets, const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags,
                                    uint8_t *ok);
int bn254_verify_batch_randomized_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets,
                                      const uint8_t *sigs, size_t n, const uint8_t *dst, size_t dst_len, uint32_t flags,
                                      const uint8_t *seed, int rand_bits, size_t group, uint8_t *ok_groups);
int bn254_aggregate_partial_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *partial,
                                uint8_t *bad);

int bn254_aggregate_verify_dist_dev(bn254_ctx *ctx, int variant, const uint8_t *pks, const uint8_t *msgs, const uint64_t *offsets, size_t n,
                                    const uint8_t *agg_sig, const uint8_t *dst, size_t dst_len, uint32_t flags, uint8_t *ok /* device byte */);

/* ---- INT32 multiply-add micro-benchmark: the roofline denominator.  Runs `iters` rounds of
 *      independent multiply-add chains on every SM of the first device (8 chains per thread, full
 *      occupancy) and returns the sustained rate in IMAD-ops/s.  mode 0: mad.lo.u32; mode 1:
 *      mad.lo.u32 + mad.hi.u32 pairs; mode 2: 32x32+64 wide multiply-add (IMAD.WIDE.U32, the
 *      instruction the field multiplication issues), counted as 2 IMAD-ops each. ------------------ */
int bn254_imad_peak(bn254_ctx *ctx, int mode, int iters, double *ops_per_s, double *ms);
/* Field-level speed of light: `warps_per_sm` warps per SM of dependent fused dot products (mode 0) or
 * Montgomery products (mode 1) on registers; canonical IMAD-ops/s (264 per product, 528 per dot product). */
int bn254_fp_peak(bn254_ctx *ctx, int mode, int warps_per_sm, int iters, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
