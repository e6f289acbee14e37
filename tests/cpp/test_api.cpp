// C++ host-side test over include/bls_bn254_b200.hpp, written the way the reference's own unit tests read
// (g1.rs:993-1140 hash / encode vectors, g2.rs:1040-1312, pairings.rs:970-980 Gt::generator, codec round trips g1.rs:937-965),
// plus the scheme layer.  Vectors come from tests/golden/reference_kats.json through a "name hex" file written by the
// pytest wrapper (tests/test_cpp_api.py).  Exit code: 0 all passed, 1 a check failed, 77 no CUDA device (skipped).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "bls_bn254_b200.hpp"

using namespace bls_bn254;

static int failures = 0;
#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                  \
        }                                                                \
    } while (0)

static Bytes unhex(const std::string &h) {
    Bytes b(h.size() / 2);
    for (size_t i = 0; i < b.size(); i++) b[i] = static_cast<uint8_t>(std::stoi(h.substr(2 * i, 2), nullptr, 16));
    return b;
}
static Bytes str(const std::string &s) { return Bytes(s.begin(), s.end()); }

int main(int argc, char **argv) {
    if (argc < 2) {
        std::printf("usage: test_api KATFILE\n");
        return 2;
    }
    std::multimap<std::string, std::vector<std::string> > kat;
    std::ifstream in(argv[1]);
    for (std::string line; std::getline(in, line);) {
        std::istringstream ss(line);
        std::string name, tok;
        ss >> name;
        std::vector<std::string> f;
        while (ss >> tok) f.push_back(tok == "-" ? "" : tok);
        kat.insert(std::make_pair(name, f));
    }
    Context *pctx = nullptr;
    try {
        pctx = new Context();
    } catch (const Bn254Error &e) {
        if (e.kind() == Bn254Error::NoDevice) {
            std::printf("no CUDA device: %s\n", e.what());
            return 77;
        }
        throw;
    }
    Context &ctx = *pctx;

    // hash_to_curve / encode_to_curve vectors: "g1_hash msghex dsthex expected(64 B)" ...
    int nvec = 0;
    for (auto &kv : kat) {
        const std::vector<std::string> &f = kv.second;
        if (kv.first == "g1_hash") { G1Affine p = G1Affine::hash(ctx, unhex(f[0]), unhex(f[1])); CHECK(Bytes(p.bytes.begin(), p.bytes.end()) == unhex(f[2])); CHECK(p.is_torsion_free(ctx)); nvec++; }
        if (kv.first == "g1_encode") { G1Affine p = G1Affine::encode(ctx, unhex(f[0]), unhex(f[1])); CHECK(Bytes(p.bytes.begin(), p.bytes.end()) == unhex(f[2])); nvec++; }
        if (kv.first == "g2_hash") { G2Affine p = G2Affine::hash(ctx, unhex(f[0]), unhex(f[1])); CHECK(Bytes(p.bytes.begin(), p.bytes.end()) == unhex(f[2])); CHECK(p.is_torsion_free(ctx)); nvec++; }
        if (kv.first == "g2_encode") { G2Affine p = G2Affine::encode(ctx, unhex(f[0]), unhex(f[1])); CHECK(Bytes(p.bytes.begin(), p.bytes.end()) == unhex(f[2])); nvec++; }
    }
    CHECK(nvec == 20);

    // pairings.rs:970-980: pairing(g1, g2) == Gt::generator()
    G1Affine g1 = G1Affine::generator();
    G2Affine g2 = G2Affine::generator();
    Gt e = pairing(ctx, g1, g2);
    CHECK(Bytes(e.repr.begin(), e.repr.end()) == unhex(kat.find("gt_generator")->second[0]));
    CHECK(e != Gt::identity());
    CHECK(pairing(ctx, G1Affine::identity(), g2) == Gt::identity());   // pairings.rs:789-800
    // bilinearity through multi_miller_loop: e(aP, Q) * e(-P, aQ) == 1
    std::array<uint8_t, 32> a = {};
    a[31] = 0x2a; a[20] = 0x07;
    G1Affine aP = g1.mul(ctx, a);
    G2Affine aQ = g2.mul(ctx, a);
    G1Affine negP = g1;
    {   // -P = (x, p - y) with y = 2
        Bytes p = unhex("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd45");
        std::memcpy(&negP.bytes[32], p.data(), 32);
    }
    std::vector<std::pair<G1Affine, G2Affine> > terms;
    terms.push_back(std::make_pair(aP, g2));
    terms.push_back(std::make_pair(negP, aQ));
    CHECK(multi_miller_loop(ctx, terms).final_exponentiation(ctx) == Gt::identity());
    CHECK(pairing(ctx, aP, g2) == pairing(ctx, g1, aQ));

    // codecs: generator round trips (g1.rs:937-965, g2.rs:962-991) and hashed points
    CHECK(G1Affine::from_compressed(ctx, g1.to_compressed(ctx).data()) == g1);
    CHECK(G2Affine::from_compressed(ctx, g2.to_compressed(ctx).data()) == g2);
    G1Affine h = G1Affine::hash(ctx, str("round trip"), dst_of(MinSig));
    CHECK(G1Affine::from_compressed(ctx, h.to_compressed(ctx).data()) == h);
    CHECK(G1Affine::from_uncompressed(ctx, h.to_uncompressed().data()) == h);
    CHECK(h.is_on_curve(ctx) && h.is_torsion_free(ctx) && !h.is_identity());
    {
        G1Affine off = h;
        off.bytes[40] ^= 1;
        bool threw = false;
        try { G1Affine::from_uncompressed(ctx, off.bytes.data()); } catch (const Bn254Error &err) { threw = err.kind() == Bn254Error::InvalidG1Bytes; }
        CHECK(threw);
    }

    // scheme: sign / verify / aggregate / aggregate_verify, batch and randomised batch
    const size_t n = 37;
    std::vector<PublicKey> pks;
    std::vector<Signature> sigs;
    Messages msgs;
    for (size_t i = 0; i < n; i++) {
        SecretKey sk;
        sk.be.fill(0);
        sk.be[31] = static_cast<uint8_t>(i + 1);
        sk.be[17] = static_cast<uint8_t>(3 * i + 5);
        Bytes m = str("message #" + std::to_string(i));
        pks.push_back(sk.public_key(ctx));
        sigs.push_back(sk.sign(ctx, m));
        msgs.push(m);
    }
    CHECK(verify(ctx, pks[0], str("message #0"), sigs[0]));
    CHECK(!verify(ctx, pks[0], str("message #1"), sigs[0]));
    std::vector<uint8_t> ok = verify_batch(ctx, pks, msgs, sigs, BN254_FLAG_CHECK_SUBGROUP);
    CHECK(ok == std::vector<uint8_t>(n, 1));
    std::vector<Signature> bad = sigs;
    std::swap(bad[5], bad[6]);
    ok = verify_batch(ctx, pks, msgs, bad);
    for (size_t i = 0; i < n; i++) CHECK(ok[i] == (i == 5 || i == 6 ? 0 : 1));
    std::array<uint8_t, 32> seed;
    for (int i = 0; i < 32; i++) seed[i] = static_cast<uint8_t>(0xa0 + i);
    CHECK(verify_batch_randomized(ctx, pks, msgs, sigs, seed, 10) == std::vector<uint8_t>(4, 1));
    std::vector<uint8_t> grp = verify_batch_randomized(ctx, pks, msgs, bad, seed, 10);
    CHECK(grp.size() == 4 && grp[0] == 0 && grp[1] == 1 && grp[2] == 1 && grp[3] == 1);
    {   // prepared keys: the same verdicts with the public keys' line tables computed once
        G2Prepared prep(ctx, pks, BN254_FLAG_CHECK_SUBGROUP);
        std::vector<uint32_t> idx(pks.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = (uint32_t)i;
        CHECK(prep.size() == pks.size());
        CHECK(prep.verify_batch(ctx, idx, msgs, sigs) == verify_batch(ctx, pks, msgs, sigs));
        CHECK(prep.verify_batch(ctx, idx, msgs, bad) == verify_batch(ctx, pks, msgs, bad));
    }
    {   // same-key grouping: key j signs messages 3j .. 3j+2; one flag per key; one wrong signature rejects only its key
        std::vector<PublicKey> kpk;
        std::vector<Signature> ksig;
        Messages kmsg;
        std::vector<size_t> counts;
        for (size_t j = 0; j < 4; j++) {
            SecretKey sk;
            sk.be.fill(0);
            sk.be[31] = static_cast<uint8_t>(j + 9);
            kpk.push_back(sk.public_key(ctx));
            counts.push_back(3);
            for (size_t q = 0; q < 3; q++) {
                Bytes m = str("stream " + std::to_string(3 * j + q));
                ksig.push_back(sk.sign(ctx, m));
                kmsg.push(m);
            }
        }
        CHECK(verify_same_key_randomized(ctx, kpk, counts, kmsg, ksig, seed) == std::vector<uint8_t>(4, 1));
        std::swap(ksig[7], ksig[8]);   // both under key 2
        std::vector<uint8_t> kk = verify_same_key_randomized(ctx, kpk, counts, kmsg, ksig, seed, 128, BN254_FLAG_CHECK_SUBGROUP);
        CHECK(kk.size() == 4 && kk[0] == 1 && kk[1] == 1 && kk[2] == 0 && kk[3] == 1);
    }
    Signature agg = aggregate(ctx, sigs);
    CHECK(aggregate_verify(ctx, pks, msgs, agg));
    CHECK(!aggregate_verify(ctx, pks, msgs, sigs[0]));
    CHECK(ctx.launch_count() > 0);

    delete pctx;
    std::printf(failures ? "%d check(s) failed\n" : "all checks passed (%d)\n", failures ? failures : nvec);
    return failures ? 1 : 0;
}
