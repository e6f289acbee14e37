// bls_bn254_b200.hpp -- C++ host side over the C ABI (bls_bn254_b200.h), mirroring the reference crate's API for the
// hot path.  The reference is Rust (no toolchain in this image), so this header plays the role its `inner_types` and scheme
// types would play after the change shown in INTEGRATION.md: same names, same byte formats, same error variants.
//
//   reference (file:line under /root/reference/src)                     here
//   inner_types/g1.rs:910-928   G1Projective::hash / ::encode          G1Affine::hash / ::encode  (+ hash_batch)
//   inner_types/g2.rs:919-936   G2Projective::hash / ::encode          G2Affine::hash / ::encode  (+ hash_batch)
//   inner_types/g1.rs:283-360   to/from_compressed, to/from_uncompressed   same names on G1Affine / G2Affine
//   inner_types/g1.rs:383-391, 852-855   is_on_curve, is_torsion_free  same names
//   inner_types/pairings.rs:760-802   pairing(p, q) -> Gt             pairing(ctx, p, q)
//   inner_types/pairings.rs:808-857   multi_miller_loop(terms)        multi_miller_loop(ctx, terms) -> MillerLoopResult
//   inner_types/pairings.rs:50-178    MillerLoopResult::final_exponentiation   same name
//   inner_types/pairings.rs:499-579   Gt::to_repr / from_repr         Gt::to_repr / from_repr, ==, identity()
//   error.rs:5-10               Bn254Error variants                    Bn254Error::Kind
//   (absent upstream)  sign / verify / aggregate / aggregate_verify, batch and randomised batch forms
//
// Everything runs on the GPU through the library; there is no CPU path (Context() throws NoDevice without CUDA).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bls_bn254_b200.h"

namespace bls_bn254 {

class Bn254Error : public std::runtime_error {
public:
    enum Kind { InvalidScalarBytes = -1, InvalidG1Bytes = -2, InvalidG2Bytes = -3, InvalidGtBytes = -4, BadArgument = -10, NoDevice = -20, Cuda = -100 };
    Bn254Error(int code, const std::string &what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }
    Kind kind() const { return code_ <= -100 ? Cuda : static_cast<Kind>(code_); }

private:
    int code_;
};

typedef std::vector<uint8_t> Bytes;

// bn254_ctx: streams and scratch on one or more GPUs; single caller
class Context {
public:
    explicit Context(const std::vector<int> &devices = std::vector<int>(1, 0)) : h_(nullptr) {
        int rc = bn254_ctx_create(&h_, devices.data(), static_cast<int>(devices.size()));
        if (rc) throw Bn254Error(rc, rc == BN254_ERR_NO_DEVICE ? "no usable CUDA device (bls_bn254_b200 has no CPU path)" : "bn254_ctx_create failed");
    }
    ~Context() { bn254_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    bn254_ctx *raw() const { return h_; }
    uint64_t launch_count() const { return bn254_launch_count(h_); }
    void check(int rc) const {
        if (rc) throw Bn254Error(rc, bn254_last_error(h_));
    }

private:
    bn254_ctx *h_;
};

// messages as the C ABI wants them: one blob + offsets
struct Messages {
    Bytes blob;
    std::vector<uint64_t> offsets;
    Messages() : offsets(1, 0) {}
    explicit Messages(const std::vector<Bytes> &msgs) : offsets(1, 0) {
        for (size_t i = 0; i < msgs.size(); i++) push(msgs[i]);
    }
    void push(const Bytes &m) {
        blob.insert(blob.end(), m.begin(), m.end());
        offsets.push_back(blob.size());
    }
    size_t size() const { return offsets.size() - 1; }
    const uint8_t *data() const { return blob.empty() ? reinterpret_cast<const uint8_t *>("") : blob.data(); }
};

enum Variant { MinSig = BN254_MIN_SIG, MinPk = BN254_MIN_PK };
static const char *const DST_MIN_SIG = "BLS_SIG_BN254G1_XMD:SHA-256_SVDW_RO_NUL_";
static const char *const DST_MIN_PK = "BLS_SIG_BN254G2_XMD:SHA-256_SVDW_RO_NUL_";
inline Bytes dst_of(Variant v) {
    const char *s = v == MinSig ? DST_MIN_SIG : DST_MIN_PK;
    return Bytes(s, s + std::strlen(s));
}

struct G2Affine;

// g1.rs: uncompressed x || y, identity = (0, 1)
struct G1Affine {
    std::array<uint8_t, 64> bytes;
    static G1Affine generator() {
        G1Affine g;
        g.bytes.fill(0);
        g.bytes[31] = 1;
        g.bytes[63] = 2;
        return g;
    }
    static G1Affine identity() {
        G1Affine g;
        g.bytes.fill(0);
        g.bytes[63] = 1;
        return g;
    }
    bool is_identity() const { return *this == identity(); }
    bool operator==(const G1Affine &o) const { return bytes == o.bytes; }
    const std::array<uint8_t, 64> &to_uncompressed() const { return bytes; }
    static G1Affine from_uncompressed(const Context &ctx, const uint8_t in[64]) {  // g1.rs:336-358: decodable and on the curve
        G1Affine g;
        std::memcpy(g.bytes.data(), in, 64);
        if (!g.is_on_curve(ctx)) throw Bn254Error(BN254_ERR_INVALID_G1_BYTES, "InvalidG1Bytes");
        return g;
    }
    std::array<uint8_t, 32> to_compressed(const Context &ctx) const {
        std::array<uint8_t, 32> out;
        ctx.check(bn254_g1_compress(ctx.raw(), bytes.data(), 1, out.data()));
        return out;
    }
    static G1Affine from_compressed(const Context &ctx, const uint8_t in[32]) {
        G1Affine g;
        uint8_t st = 0;
        ctx.check(bn254_g1_decompress(ctx.raw(), in, 1, g.bytes.data(), &st));
        if (!st) throw Bn254Error(BN254_ERR_INVALID_G1_BYTES, "InvalidG1Bytes");
        return g;
    }
    bool is_on_curve(const Context &ctx) const {
        uint8_t ok = 0;
        ctx.check(bn254_g1_is_on_curve(ctx.raw(), bytes.data(), 1, &ok));
        return ok != 0;
    }
    bool is_torsion_free(const Context &ctx) const {
        uint8_t ok = 0;
        ctx.check(bn254_g1_is_torsion_free(ctx.raw(), bytes.data(), 1, &ok));
        return ok != 0;
    }
    // G1Projective::hash::<ExpandMsgXmd<Sha256>>(msg, dst) / ::encode, as an affine point
    static G1Affine hash(const Context &ctx, const Bytes &msg, const Bytes &dst) { return map(ctx, msg, dst, true); }
    static G1Affine encode(const Context &ctx, const Bytes &msg, const Bytes &dst) { return map(ctx, msg, dst, false); }
    static std::vector<G1Affine> hash_batch(const Context &ctx, const Messages &m, const Bytes &dst) {
        std::vector<G1Affine> out(m.size());
        std::vector<uint8_t> st(m.size() + 1);
        ctx.check(bn254_hash_to_g1(ctx.raw(), m.data(), m.offsets.data(), m.size(), dst.data(), dst.size(),
                                   out.empty() ? nullptr : out[0].bytes.data(), st.data()));
        for (size_t i = 0; i < m.size(); i++)
            if (!st[i]) throw Bn254Error(BN254_ERR_INVALID_G1_BYTES, "SvdW exceptional input (the reference panics here, fp.rs:325)");
        return out;
    }
    G1Affine mul(const Context &ctx, const std::array<uint8_t, 32> &scalar_be) const {  // g1.rs:821-841
        G1Affine r;
        ctx.check(bn254_g1_mul(ctx.raw(), bytes.data(), scalar_be.data(), 1, r.bytes.data()));
        return r;
    }
    static G1Affine sum(const Context &ctx, const std::vector<G1Affine> &pts) {
        G1Affine r;
        ctx.check(bn254_g1_sum(ctx.raw(), pts.empty() ? nullptr : pts[0].bytes.data(), pts.size(), r.bytes.data()));
        return r;
    }

private:
    static G1Affine map(const Context &ctx, const Bytes &msg, const Bytes &dst, bool ro) {
        G1Affine g;
        uint64_t off[2] = {0, msg.size()};
        uint8_t st = 0;
        const uint8_t *p = msg.empty() ? reinterpret_cast<const uint8_t *>("") : msg.data();
        ctx.check((ro ? bn254_hash_to_g1 : bn254_encode_to_g1)(ctx.raw(), p, off, 1, dst.data(), dst.size(), g.bytes.data(), &st));
        if (!st) throw Bn254Error(BN254_ERR_INVALID_G1_BYTES, "SvdW exceptional input (the reference panics here, fp.rs:325)");
        return g;
    }
};
static_assert(sizeof(G1Affine) == 64, "G1Affine must be the 64 wire bytes");

// g2.rs: x.c1 || x.c0 || y.c1 || y.c0
struct G2Affine {
    std::array<uint8_t, 128> bytes;
    static G2Affine generator() {  // the EIP-197 generator decoded from fp2.rs:305-348
        static const char *hex =
            "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
            "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
            "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"
            "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa";
        G2Affine g;
        for (int i = 0; i < 128; i++) {
            auto nib = [](char c) { return c <= '9' ? c - '0' : c - 'a' + 10; };
            g.bytes[i] = static_cast<uint8_t>(nib(hex[2 * i]) * 16 + nib(hex[2 * i + 1]));
        }
        return g;
    }
    static G2Affine identity() {
        G2Affine g;
        g.bytes.fill(0);
        g.bytes[127] = 1;
        return g;
    }
    bool is_identity() const { return *this == identity(); }
    bool operator==(const G2Affine &o) const { return bytes == o.bytes; }
    const std::array<uint8_t, 128> &to_uncompressed() const { return bytes; }
    static G2Affine from_uncompressed(const Context &ctx, const uint8_t in[128]) {
        G2Affine g;
        std::memcpy(g.bytes.data(), in, 128);
        if (!g.is_on_curve(ctx)) throw Bn254Error(BN254_ERR_INVALID_G2_BYTES, "InvalidG2Bytes");
        return g;
    }
    std::array<uint8_t, 64> to_compressed(const Context &ctx) const {
        std::array<uint8_t, 64> out;
        ctx.check(bn254_g2_compress(ctx.raw(), bytes.data(), 1, out.data()));
        return out;
    }
    static G2Affine from_compressed(const Context &ctx, const uint8_t in[64]) {
        G2Affine g;
        uint8_t st = 0;
        ctx.check(bn254_g2_decompress(ctx.raw(), in, 1, g.bytes.data(), &st));
        if (!st) throw Bn254Error(BN254_ERR_INVALID_G2_BYTES, "InvalidG2Bytes");
        return g;
    }
    bool is_on_curve(const Context &ctx) const {
        uint8_t ok = 0;
        ctx.check(bn254_g2_is_on_curve(ctx.raw(), bytes.data(), 1, &ok));
        return ok != 0;
    }
    bool is_torsion_free(const Context &ctx) const {
        uint8_t ok = 0;
        ctx.check(bn254_g2_is_torsion_free(ctx.raw(), bytes.data(), 1, &ok));
        return ok != 0;
    }
    static G2Affine hash(const Context &ctx, const Bytes &msg, const Bytes &dst) { return map(ctx, msg, dst, true); }
    static G2Affine encode(const Context &ctx, const Bytes &msg, const Bytes &dst) { return map(ctx, msg, dst, false); }
    G2Affine mul(const Context &ctx, const std::array<uint8_t, 32> &scalar_be) const {  // g2.rs:866-886
        G2Affine r;
        ctx.check(bn254_g2_mul(ctx.raw(), bytes.data(), scalar_be.data(), 1, r.bytes.data()));
        return r;
    }
    static G2Affine sum(const Context &ctx, const std::vector<G2Affine> &pts) {
        G2Affine r;
        ctx.check(bn254_g2_sum(ctx.raw(), pts.empty() ? nullptr : pts[0].bytes.data(), pts.size(), r.bytes.data()));
        return r;
    }

private:
    static G2Affine map(const Context &ctx, const Bytes &msg, const Bytes &dst, bool ro) {
        G2Affine g;
        uint64_t off[2] = {0, msg.size()};
        uint8_t st = 0;
        const uint8_t *p = msg.empty() ? reinterpret_cast<const uint8_t *>("") : msg.data();
        ctx.check((ro ? bn254_hash_to_g2 : bn254_encode_to_g2)(ctx.raw(), p, off, 1, dst.data(), dst.size(), g.bytes.data(), &st));
        if (!st) throw Bn254Error(BN254_ERR_INVALID_G2_BYTES, "SvdW exceptional input (the reference panics here, fp2.rs:255)");
        return g;
    }
};
static_assert(sizeof(G2Affine) == 128, "G2Affine must be the 128 wire bytes");

// pairings.rs:221-249, 491-579
struct Gt {
    std::array<uint8_t, 384> repr;
    static Gt identity() {
        Gt g;
        g.repr.fill(0);
        g.repr[31] = 1;
        return g;
    }
    bool operator==(const Gt &o) const { return repr == o.repr; }
    bool operator!=(const Gt &o) const { return !(*this == o); }
    const std::array<uint8_t, 384> &to_repr() const { return repr; }
    static Gt from_repr(const uint8_t in[384]) {
        Gt g;
        std::memcpy(g.repr.data(), in, 384);
        return g;
    }
};

// pairings.rs:20-43
struct MillerLoopResult {
    std::array<uint8_t, 384> repr;
    Gt final_exponentiation(const Context &ctx) const {  // pairings.rs:50-178
        Gt g;
        uint8_t st = 0;
        ctx.check(bn254_final_exponentiation(ctx.raw(), repr.data(), 1, g.repr.data(), &st));
        if (!st) throw Bn254Error(BN254_ERR_INVALID_GT_BYTES, "InvalidGtBytes");
        return g;
    }
};

// pairings.rs:760-802
inline Gt pairing(const Context &ctx, const G1Affine &p, const G2Affine &q) {
    Gt g;
    uint8_t st = 0;
    ctx.check(bn254_pairing_batch(ctx.raw(), p.bytes.data(), q.bytes.data(), 1, g.repr.data(), &st));
    if (!st) throw Bn254Error(BN254_ERR_INVALID_G1_BYTES, "undecodable pairing input");
    return g;
}
// pairings.rs:808-857: product of the Miller loops of all terms (identity terms contribute 1)
inline MillerLoopResult multi_miller_loop(const Context &ctx, const std::vector<std::pair<G1Affine, G2Affine> > &terms) {
    Bytes g1(64 * terms.size()), g2(128 * terms.size());
    for (size_t i = 0; i < terms.size(); i++) {
        std::memcpy(&g1[64 * i], terms[i].first.bytes.data(), 64);
        std::memcpy(&g2[128 * i], terms[i].second.bytes.data(), 128);
    }
    MillerLoopResult r;
    ctx.check(bn254_multi_miller_loop(ctx.raw(), g1.data(), g2.data(), terms.size(), r.repr.data()));
    return r;
}

// ---- the scheme layer (absent in the reference snapshot), min-sig: keys in G2, signatures in G1 --------
struct SecretKey {
    std::array<uint8_t, 32> be;  // scalar below r, big-endian
    G2Affine public_key(const Context &ctx) const { return G2Affine::generator().mul(ctx, be); }
    G1Affine sign(const Context &ctx, const Bytes &msg) const { return G1Affine::hash(ctx, msg, dst_of(MinSig)).mul(ctx, be); }
};
typedef G2Affine PublicKey;
typedef G1Affine Signature;

// n independent (pk, msg, sig) tuples -> n accept flags (bn254_verify_batch)
inline std::vector<uint8_t> verify_batch(const Context &ctx, const std::vector<PublicKey> &pks, const Messages &msgs, const std::vector<Signature> &sigs,
                                         uint32_t flags = 0) {
    size_t n = msgs.size();
    if (pks.size() != n || sigs.size() != n) throw Bn254Error(BN254_ERR_ARG, "length mismatch");
    std::vector<uint8_t> ok(n);
    Bytes dst = dst_of(MinSig);
    if (n)
        ctx.check(bn254_verify_batch(ctx.raw(), BN254_MIN_SIG, pks[0].bytes.data(), msgs.data(), msgs.offsets.data(), sigs[0].bytes.data(), n, dst.data(),
                                     dst.size(), flags, ok.data()));
    return ok;
}
inline bool verify(const Context &ctx, const PublicKey &pk, const Bytes &msg, const Signature &sig, uint32_t flags = 0) {
    Messages m;
    m.push(msg);
    return verify_batch(ctx, std::vector<PublicKey>(1, pk), m, std::vector<Signature>(1, sig), flags)[0] != 0;
}
// one accept flag per group of `group` tuples (bn254_verify_batch_randomized); `seed` = 32 bytes the signers cannot predict
inline std::vector<uint8_t> verify_batch_randomized(const Context &ctx, const std::vector<PublicKey> &pks, const Messages &msgs,
                                                    const std::vector<Signature> &sigs, const std::array<uint8_t, 32> &seed, size_t group = 256,
                                                    int rand_bits = 64, uint32_t flags = 0) {
    size_t n = msgs.size();
    if (pks.size() != n || sigs.size() != n) throw Bn254Error(BN254_ERR_ARG, "length mismatch");
    std::vector<uint8_t> ok(group ? (n + group - 1) / group : 0);
    Bytes dst = dst_of(MinSig);
    if (n)
        ctx.check(bn254_verify_batch_randomized(ctx.raw(), BN254_MIN_SIG, pks[0].bytes.data(), msgs.data(), msgs.offsets.data(), sigs[0].bytes.data(), n,
                                                dst.data(), dst.size(), flags, seed.data(), rand_bits, group, ok.data()));
    return ok;
}
// Same-key randomised verification (bn254_verify_same_key_randomized): key k signed the next msg_counts[k] messages; one accept
// flag per KEY from one Miller loop and one final exponentiation over the randomised sums of its signatures and message points.
inline std::vector<uint8_t> verify_same_key_randomized(const Context &ctx, const std::vector<PublicKey> &pks, const std::vector<size_t> &msg_counts,
                                                       const Messages &msgs, const std::vector<Signature> &sigs, const std::array<uint8_t, 32> &seed,
                                                       int rand_bits = 64, uint32_t flags = 0) {
    const size_t nk = pks.size(), n = msgs.size();
    std::vector<uint64_t> koff(nk + 1, 0);
    if (msg_counts.size() != nk || sigs.size() != n) throw Bn254Error(BN254_ERR_ARG, "length mismatch");
    for (size_t k = 0; k < nk; k++) koff[k + 1] = koff[k] + msg_counts[k];
    if (koff[nk] != n) throw Bn254Error(BN254_ERR_ARG, "msg_counts must add up to the number of messages");
    std::vector<uint8_t> ok(nk);
    Bytes dst = dst_of(MinSig);
    if (nk)
        ctx.check(bn254_verify_same_key_randomized(ctx.raw(), BN254_MIN_SIG, pks[0].bytes.data(), koff.data(), nk, msgs.data(), msgs.offsets.data(),
                                                   n ? sigs[0].bytes.data() : nullptr, dst.data(), dst.size(), flags, seed.data(), rand_bits, ok.data()));
    return ok;
}
// G2Prepared (pairings.rs:609-660) for a set of public keys: prepare once, verify many.  Tuple i of verify_batch names key
// key_index[i]; the accept flags equal verify_batch's on the expanded key array.
class G2Prepared {
  public:
    G2Prepared(const Context &ctx, const std::vector<PublicKey> &keys, uint32_t flags = 0) : h_(nullptr) {
        ctx.check(bn254_g2_prepare(ctx.raw(), keys.empty() ? nullptr : keys[0].bytes.data(), keys.size(), flags, &h_));
    }
    ~G2Prepared() { bn254_g2_prepared_free(h_); }
    G2Prepared(const G2Prepared &) = delete;
    G2Prepared &operator=(const G2Prepared &) = delete;
    size_t size() const { return bn254_g2_prepared_count(h_); }
    std::vector<uint8_t> verify_batch(const Context &ctx, const std::vector<uint32_t> &key_index, const Messages &msgs, const std::vector<Signature> &sigs,
                                      uint32_t flags = 0) const {
        size_t n = msgs.size();
        if (key_index.size() != n || sigs.size() != n) throw Bn254Error(BN254_ERR_ARG, "length mismatch");
        std::vector<uint8_t> ok(n);
        Bytes dst = dst_of(MinSig);
        if (n)
            ctx.check(bn254_verify_batch_prepared(ctx.raw(), h_, key_index.data(), msgs.data(), msgs.offsets.data(), sigs[0].bytes.data(), n, dst.data(),
                                                  dst.size(), flags, ok.data()));
        return ok;
    }

  private:
    bn254_g2_prepared *h_;
};
inline Signature aggregate(const Context &ctx, const std::vector<Signature> &sigs) { return G1Affine::sum(ctx, sigs); }
// distinct messages: e(agg, -g2) * prod e(H(m_i), pk_i) == 1
inline bool aggregate_verify(const Context &ctx, const std::vector<PublicKey> &pks, const Messages &msgs, const Signature &agg, uint32_t flags = 0) {
    size_t n = msgs.size();
    if (pks.size() != n) throw Bn254Error(BN254_ERR_ARG, "length mismatch");
    uint8_t ok = 0;
    Bytes dst = dst_of(MinSig);
    ctx.check(bn254_aggregate_verify(ctx.raw(), BN254_MIN_SIG, n ? pks[0].bytes.data() : nullptr, msgs.data(), msgs.offsets.data(), n, agg.bytes.data(),
                                     dst.data(), dst.size(), flags, &ok));
    return ok != 0;
}

}  // namespace bls_bn254
