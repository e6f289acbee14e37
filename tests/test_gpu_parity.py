"""GPU parity suite (-m gpu): every result of libbls_bn254_b200.so, called through the C ABI, is
compared byte for byte with (1) the reference's golden vectors (tests/golden/reference_kats.json),
(2) the CPU oracle on the same seeded inputs, and (3) at BASELINE.json's sizes, size-independent
properties (accept bits known by construction, bilinearity, product-of-shards)."""
import random

import pytest

pytestmark = pytest.mark.gpu

H = bytes.fromhex
QUUX_G1 = b"QUUX-V01-CS02-with-BN254G1_XMD:SHA-256_SVDW_RO_"
QUUX_G2 = b"QUUX-V01-CS02-with-BN254G2_XMD:SHA-256_SVDW_RO_"


@pytest.fixture(scope="module")
def ctx():
    import bls_bn254_b200 as b
    c = b.Context([0])
    yield c
    c.close()


@pytest.fixture(scope="module")
def B():
    import bls_bn254_b200 as b
    return b


@pytest.fixture(params=[1, 2, 3, 4], ids=["one-thread", "lane-pair", "slot-machine", "wide"])
def shape(request, ctx):
    """Force one shape of the pairing kernels for the test (bn254_ctx_set_option "lanes"): 1 = one thread per pairing, 2 = two
    lanes per pairing (tower code on local memory), 3 = the slot machine (vm.cuh: two lanes per pairing, operands in shared
    memory, the tower as constant-memory programs), 4 = the wide slot machine (one warp per pairing, up to 16 independent
    instructions at a time: the latency shape).  Results are bit-identical."""
    ctx.set_option("lanes", request.param)
    yield request.param
    ctx.set_option("lanes", 0)


def launched(ctx, fn):
    """run fn() with per-launch profiling on; returns (result, set of kernel names launched)"""
    ctx.profile(True)
    ctx.profile_drain()
    try:
        out = fn()
        names = {r[0] for r in ctx.profile_drain()}
    finally:
        ctx.profile(False)
    return out, names


def rand_msgs(rng, n, maxlen=96):
    out = [b"", b"abc", b"abcdef0123456789", b"q128_" + b"q" * 128, b"a512_" + b"a" * 512]
    while len(out) < n:
        out.append(rng.randbytes(rng.randrange(0, maxlen)))
    return out[:n]


# ---------------------------------------------------------------------------------- golden vectors
def test_g1_kats(ctx, kats):
    for suite, fn in (("g1_encode", ctx.encode_to_g1), ("g1_hash", ctx.hash_to_g1)):
        dst = kats[suite + "_dst"].encode()
        msgs = [v["msg"].encode("latin1") for v in kats[suite]]
        out, st = fn(msgs, dst, want_status=True)
        assert st == b"\x01" * len(msgs)
        for i, v in enumerate(kats[suite]):
            assert out[64 * i:64 * i + 64] == H(v["p_x"]) + H(v["p_y"]), (suite, i)


def test_g2_kats(ctx, kats):
    for suite, fn in (("g2_encode", ctx.encode_to_g2), ("g2_hash", ctx.hash_to_g2)):
        dst = kats[suite + "_dst"].encode()
        msgs = [v["msg"].encode("latin1") for v in kats[suite]]
        out = fn(msgs, dst)
        for i, v in enumerate(kats[suite]):
            assert out[128 * i:128 * i + 128] == H(v["x_c1"]) + H(v["x_c0"]) + H(v["y_c1"]) + H(v["y_c0"]), (suite, i)


def test_gt_generator_kat(ctx, kats, B):
    """pairings.rs:970-980: pairing(g1, g2) == Gt::generator()."""
    exp = b"".join(H(x) for x in kats["gt_generator"])
    assert ctx.pairing(B.G1_GENERATOR, B.G2_GENERATOR) == exp


# -------------------------------------------------------------------------------- against the oracle
def test_hash_vs_oracle(ctx, orc):
    rng = random.Random(11)
    msgs = rand_msgs(rng, 300)
    for dst in (QUUX_G1, b"", b"x" * 255, b"y" * 256, b"z" * 700):
        sub = msgs if dst == QUUX_G1 else msgs[:40]
        assert ctx.hash_to_g1(sub, dst) == orc.hash_to_g1(sub, dst, 2, nthreads=8)
        assert ctx.encode_to_g1(sub, dst) == orc.hash_to_g1(sub, dst, 1, nthreads=8)
    sub = msgs[:120]
    assert ctx.hash_to_g2(sub, QUUX_G2) == orc.hash_to_g2(sub, QUUX_G2, 2, nthreads=8)
    assert ctx.encode_to_g2(sub, QUUX_G2) == orc.hash_to_g2(sub, QUUX_G2, 1, nthreads=8)
    assert ctx.hash_to_g2(sub[:20], b"w" * 300) == orc.hash_to_g2(sub[:20], b"w" * 300, 2, nthreads=8)


def test_hash_empty_batch(ctx):
    assert ctx.hash_to_g1([], QUUX_G1) == b""
    assert ctx.hash_to_g2([], QUUX_G2) == b""


def _points(ctx, orc, n, seed):
    rng = random.Random(seed)
    msgs = [rng.randbytes(16) for _ in range(n)]
    return orc.hash_to_g1(msgs, QUUX_G1, nthreads=8), orc.hash_to_g2(msgs, QUUX_G2, nthreads=8)


def test_pairing_vs_oracle(ctx, orc, B, shape):
    g1, g2 = _points(ctx, orc, 40, 3)
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    ident2 = bytes(64) + bytes(32) + bytes(31) + b"\x01"
    bad1 = b"\xff" * 64  # x >= p after masking
    bad2 = b"\xff" * 128
    g1 = B.G1_GENERATOR + ident1 + g1[:64] + bad1 + g1
    g2 = B.G2_GENERATOR + g2[:128] + ident2 + g2[:128] + g2
    # one more undecodable G2
    g1 += g1[-64:]
    g2 += bad2
    gt, st = ctx.pairing(g1, g2, want_status=True)
    egt, est = orc.pairing_batch(g1, g2, nthreads=8, want_status=True)
    assert st == est
    assert gt == egt
    ml = ctx.miller_loop(g1, g2)
    assert ml == orc.miller_loop_batch(g1, g2, nthreads=8)
    # final exponentiation of the valid Miller-loop results
    good = b"".join(ml[384 * i:384 * i + 384] for i in range(len(st)) if st[i])
    assert ctx.final_exponentiation(good) == orc.final_exponentiation(good, nthreads=8)


def test_multi_miller_and_product(ctx, orc, B, shape):
    g1, g2 = _points(ctx, orc, 37, 5)
    assert ctx.multi_miller_loop(g1, g2) == orc.multi_miller_loop(g1, g2)
    assert ctx.multi_miller_loop(b"", b"") == orc.multi_miller_loop(b"", b"")
    ml = orc.miller_loop_batch(g1, g2, nthreads=8)
    assert ctx.fp12_product(ml) == orc.fp12_product(ml)
    assert ctx.fp12_product(ml[:384]) == ml[:384]
    with pytest.raises(B.Bn254Error) as e:
        ctx.multi_miller_loop(b"\xff" * 64, g2[:128])
    assert e.value.code == -2
    with pytest.raises(B.Bn254Error) as e:
        ctx.fp12_product(b"\xff" * 384)
    assert e.value.code == -4


def test_bilinearity(ctx, orc):
    """e(aP, bQ) == e(abP, Q) == e(P, abQ) on the GPU path alone."""
    g1, g2 = _points(ctx, orc, 4, 7)
    a = (123456789123456789 << 64 | 987654321).to_bytes(32, "big")
    b = (0xdeadbeefcafebabe1122334455667788).to_bytes(32, "big")
    ab = ((int.from_bytes(a, "big") * int.from_bytes(b, "big")) % 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001).to_bytes(32, "big")
    aP = ctx.g1_mul(g1, a * 4)
    bQ = ctx.g2_mul(g2, b * 4)
    abP = ctx.g1_mul(g1, ab * 4)
    abQ = ctx.g2_mul(g2, ab * 4)
    x = ctx.pairing(aP, bQ)
    assert x == ctx.pairing(abP, g2) == ctx.pairing(g1, abQ)


def test_group_ops_vs_oracle(ctx, orc, B):
    rng = random.Random(9)
    g1, g2 = _points(ctx, orc, 24, 9)
    sc = b"".join(rng.randrange(1 << 254).to_bytes(32, "big") for _ in range(24))
    sc = (0).to_bytes(32, "big") + (1).to_bytes(32, "big") + sc[64:]
    assert ctx.g1_mul(g1, sc) == orc.g1_mul(g1, sc, nthreads=8)
    assert ctx.g2_mul(g2, sc) == orc.g2_mul(g2, sc, nthreads=8)
    assert ctx.g1_sum(g1) == orc.g1_sum(g1)
    assert ctx.g2_sum(g2) == orc.g2_sum(g2)
    assert ctx.g1_sum(g1[:64]) == g1[:64]
    # 2*2*G + G == 5*G (g1.rs:1143-1149, g2.rs:1030-1036)
    five = (5).to_bytes(32, "big")
    four = (4).to_bytes(32, "big")
    assert ctx.g1_sum(ctx.g1_mul(B.G1_GENERATOR, four) + B.G1_GENERATOR) == ctx.g1_mul(B.G1_GENERATOR, five)
    assert ctx.g2_sum(ctx.g2_mul(B.G2_GENERATOR, four) + B.G2_GENERATOR) == ctx.g2_mul(B.G2_GENERATOR, five)


def test_codecs_vs_oracle(ctx, orc, B):
    g1, g2 = _points(ctx, orc, 50, 13)
    ident1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    g1 = B.G1_GENERATOR + ident1 + g1
    g2 = B.G2_GENERATOR + g2
    c1, c2 = ctx.g1_compress(g1), ctx.g2_compress(g2)
    assert c1 == orc.g1_compress(g1)
    assert c2 == orc.g2_compress(g2)
    assert c1[32:64] == b"\x80" + bytes(31)  # identity (g1.rs:265-288)
    d1, s1 = ctx.g1_decompress(c1)
    assert (d1, s1) == orc.g1_decompress(c1)
    assert d1 == g1
    d2, s2 = ctx.g2_decompress(c2)
    assert (d2, s2) == orc.g2_decompress(c2)
    assert d2 == g2
    # x with no point on the curve / non-canonical x
    junk = b"".join((i + 5).to_bytes(32, "big") for i in range(20)) + b"\x7f" + b"\xff" * 31
    assert ctx.g1_decompress(junk) == orc.g1_decompress(junk)
    junk2 = b"".join((i + 5).to_bytes(64, "big") for i in range(12))
    assert ctx.g2_decompress(junk2) == orc.g2_decompress(junk2)
    off = bytearray(g1[:64 * 6])
    off[70] ^= 1
    assert ctx.g1_is_on_curve(bytes(off)) == orc.g1_is_on_curve(bytes(off))
    off2 = bytearray(g2[:128 * 5])
    off2[300] ^= 1
    assert ctx.g2_is_on_curve(bytes(off2)) == orc.g2_is_on_curve(bytes(off2))


@pytest.mark.parametrize("variant", [0, 1])
def test_verify_vs_oracle(ctx, orc, B, variant, shape):
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, 96, variant=variant, corrupt_every=8)
    dst = B.DST_MIN_SIG if variant == 0 else B.DST_MIN_PK
    rpk, rsig = (128, 64) if variant == 0 else (64, 128)
    # add an undecodable key and an identity signature
    pks = pks + b"\xff" * rpk + pks[:rpk]
    ident = ((0).to_bytes(32, "big") + (1).to_bytes(32, "big")) if rsig == 64 else (bytes(64) + bytes(63) + b"\x01")
    sigs = sigs + sigs[:rsig] + ident
    msgs = msgs + [b"extra-1", b""]
    got = ctx.verify_batch(variant, pks, msgs, sigs, dst)
    ref = orc.verify_batch(variant, pks, msgs, sigs, dst, nthreads=8)
    assert got == ref
    assert got[:96] == exp
    assert 0 < sum(exp) < 96
    # on-curve flag
    bad = bytearray(sigs)
    bad[rsig - 1] ^= 1
    got = ctx.verify_batch(variant, pks, msgs, bytes(bad), dst, flags=B.CHECK_ON_CURVE)
    assert got == orc.verify_batch(variant, pks, msgs, bytes(bad), dst, flags=1, nthreads=8)
    assert got[0] == 0
    assert ctx.verify_batch(variant, b"", [], b"", dst) == b""
    # the kernels of the forced shape really ran
    _, names = launched(ctx, lambda: ctx.verify_batch(variant, pks, msgs, sigs, dst))
    want = {0: ("k_miller2_fixed", "k_final_exp"), 1: ("k_miller2", "k_final_exp")}[variant]
    if shape in (3, 4):
        assert {("k_vm_" if shape == 3 else "k_w_") + w[2:] for w in want} <= names
    else:
        assert {w + ("_l2" if shape == 2 else "") for w in want} <= names


@pytest.mark.parametrize("variant", [0, 1])
def test_identity_and_y_equals_one_inputs_never_verify(ctx, orc, B, variant):
    """an identity public key drops out of the Miller product, so (pk = identity, sig = identity)
    used to verify on every message; so did pk = (x, 1) / sig = (x, 1), which the reference's decode convention flags as
    infinity (g1.rs:318-324) -- an off-curve forgery under default flags.  All of them are rejected now, by the exact
    predicate, the randomised one and the aggregate one, on the GPU and in the oracle."""
    n = 12
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, n, variant=variant, corrupt_every=0)
    dst = B.DST_MIN_SIG if variant == 0 else B.DST_MIN_PK
    rpk, rsig = (128, 64) if variant == 0 else (64, 128)
    id1 = (0).to_bytes(32, "big") + (1).to_bytes(32, "big")
    id2 = bytes(64) + bytes(63) + b"\x01"
    y1_g1 = (12345).to_bytes(32, "big") + (1).to_bytes(32, "big")          # any x, y = 1: "infinity" by convention, not on the curve
    y1_g2 = (7).to_bytes(32, "big") + (9).to_bytes(32, "big") + bytes(63) + b"\x01"
    idpk, idsig = (id2, id1) if variant == 0 else (id1, id2)
    y1pk, y1sig = (y1_g2, y1_g1) if variant == 0 else (y1_g1, y1_g2)
    P, S = bytearray(pks), bytearray(sigs)
    P[0:rpk], S[0:rsig] = idpk, idsig                    # tuple 0: identity key and identity signature
    P[2 * rpk:3 * rpk], S[2 * rsig:3 * rsig] = y1pk, y1sig   # tuple 2: the y == 1 forgery
    P[4 * rpk:5 * rpk] = idpk                            # tuple 4: identity key, honest signature
    S[6 * rsig:7 * rsig] = idsig                         # tuple 6: honest key, identity signature
    P, S = bytes(P), bytes(S)
    want = bytearray(exp)
    for j in (0, 2, 4, 6):
        want[j] = 0
    for flags in (0, 1, 3):
        got = ctx.verify_batch(variant, P, msgs, S, dst, flags=flags)
        assert got == bytes(want), flags
        assert got == orc.verify_batch(variant, P, msgs, S, dst, flags=flags, nthreads=8)
    seed = bytes(range(32))
    got = ctx.verify_batch_randomized(variant, P, msgs, S, dst, seed, group=1)
    assert got == bytes(want)
    got = ctx.verify_batch_randomized(variant, P, msgs, S, dst, seed, group=4)
    assert got == b"\x00\x00\x01" == orc.verify_batch_randomized(variant, P, msgs, S, dst, seed, group=4, nthreads=8)
    # aggregate: an identity key (or the y == 1 key) rejects even though the product itself would be 1
    agg = B.scheme.aggregate(ctx, sigs, variant)
    assert ctx.aggregate_verify(variant, pks, msgs, agg, dst) == 1
    for bad_pk in (idpk, y1pk):
        Q = bad_pk + pks[rpk:]
        rest = B.scheme.aggregate(ctx, sigs[rsig:], variant)    # what the product needs once tuple 0 dropped out
        assert ctx.aggregate_verify(variant, Q, msgs, rest, dst) == 0 == orc.aggregate_verify(variant, Q, msgs, rest, dst, nthreads=8)
    # an identity AGGREGATE signature is a legal group element: the verdict is the product's
    assert ctx.aggregate_verify(variant, pks, msgs, idsig, dst) == 0 == orc.aggregate_verify(variant, pks, msgs, idsig, dst, nthreads=8)


@pytest.mark.parametrize("variant", [0, 1])
def test_aggregate_verify(ctx, orc, B, variant, shape):
    n = 50
    pks, msgs, sigs, _ = B.scheme.synth_batch(ctx, n, variant=variant, corrupt_every=0)
    dst = B.DST_MIN_SIG if variant == 0 else B.DST_MIN_PK
    agg = B.scheme.aggregate(ctx, sigs, variant)
    assert agg == (orc.g1_sum(sigs) if variant == 0 else orc.g2_sum(sigs))
    assert ctx.aggregate_verify(variant, pks, msgs, agg, dst) == 1 == orc.aggregate_verify(variant, pks, msgs, agg, dst, nthreads=8)
    msgs2 = list(msgs)
    msgs2[17] = b"tampered"
    assert ctx.aggregate_verify(variant, pks, msgs2, agg, dst) == 0 == orc.aggregate_verify(variant, pks, msgs2, agg, dst, nthreads=8)
    # sharded: three partial products + one finish == the single call; partial product == oracle's Miller product
    rpk = 128 if variant == 0 else 64
    cuts = [0, 13, 31, n]
    parts, bads = b"", []
    for s in range(3):
        lo, hi = cuts[s], cuts[s + 1]
        p, b = ctx.aggregate_partial(variant, pks[lo * rpk:hi * rpk], msgs[lo:hi], agg if s == 1 else None, dst)
        parts += p
        bads.append(b)
    assert bads == [0, 0, 0]
    assert ctx.aggregate_finish(parts, bads) == 1
    assert ctx.aggregate_finish(parts, [0, 1, 0]) == 0
    assert ctx.aggregate_finish(parts[:768], [0, 0]) == 0
    # a bad key anywhere rejects
    badpk = b"\xff" * rpk + pks[rpk:]
    assert ctx.aggregate_verify(variant, badpk, msgs, agg, dst) == 0


@pytest.mark.parametrize("variant", [0, 1])
def test_randomized_batch_verify_vs_oracle(ctx, orc, B, variant, shape):
    """bn254_verify_batch_randomized: group verdicts equal the oracle's and the AND of the exact
    per-tuple verdicts; ragged last group, group = 1, group > n, both scalar widths."""
    n = 203
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, n, variant=variant, corrupt_every=40)
    dst = B.DST_MIN_SIG if variant == 0 else B.DST_MIN_PK
    seed = bytes(range(7, 39))
    for group, bits in ((16, 64), (64, 128), (1, 64), (500, 64)):
        got = ctx.verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, group=group, rand_bits=bits)
        want = bytes(int(all(exp[g:g + group])) for g in range(0, n, group))
        assert got == want
        if group != 1:
            assert got == orc.verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, group=group, rand_bits=bits, nthreads=8)
    assert 0 < sum(ctx.verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, group=16)) < 13
    # undecodable input and the on-curve flag reject the group that holds them
    rpk, rsig = (128, 64) if variant == 0 else (64, 128)
    junk = b"\xff" * rpk + pks[rpk:]
    assert ctx.verify_batch_randomized(variant, junk, msgs, sigs, dst, seed, group=16)[0] == 0
    off = bytearray(sigs)
    off[rsig * 100 + rsig - 1] ^= 1
    got = ctx.verify_batch_randomized(variant, pks, msgs, bytes(off), dst, seed, group=16, flags=B.CHECK_ON_CURVE)
    assert got == orc.verify_batch_randomized(variant, pks, msgs, bytes(off), dst, seed, group=16, flags=1, nthreads=8)
    assert got[100 // 16] == 0
    assert ctx.verify_batch_randomized(variant, b"", [], b"", dst, seed) == b""
    # the wrapper with per-tuple verdicts: accepted groups wholesale, rejected groups re-verified exactly
    assert B.scheme.verify_randomized(ctx, pks, msgs, sigs, variant=variant, group=32) == exp
    with pytest.raises(B.Bn254Error):
        ctx.verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, group=0)
    with pytest.raises(B.Bn254Error):
        ctx.verify_batch_randomized(variant, pks, msgs, sigs, dst, seed, rand_bits=32)


def test_randomized_batch_verify_scalars_and_cancellation(ctx, orc, B):
    """Errors that cancel in a plain sum are caught; a forgery built from PREDICTED scalars is accepted, which
    pins r_i = SHA-256(seed || LE64(i))[:rand_bits/8] bit-exactly (and shows why the seed must be secret)."""
    from rbv_util import _neg_g1, rbv_forgery
    n = 40
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, n, variant=0, corrupt_every=0)
    dst, seed = B.DST_MIN_SIG, bytes(range(32))
    d = B.G1_GENERATOR
    forged = bytearray(sigs)
    forged[0:64] = ctx.g1_sum(sigs[0:64] + d)
    forged[64:128] = ctx.g1_sum(sigs[64:128] + _neg_g1(d))
    forged = bytes(forged)
    assert ctx.verify_batch(0, pks, msgs, forged, dst)[:3] == b"\x00\x00\x01"
    assert ctx.aggregate_verify(0, pks, msgs, ctx.g1_sum(forged), dst) == 1
    assert ctx.verify_batch_randomized(0, pks, msgs, forged, dst, seed, group=n) == b"\x00"
    for bits, (i, j) in ((64, (0, 1)), (128, (3, 38)), (64, (17, 33))):
        known = rbv_forgery(orc, sigs, seed, bits=bits, i=i, j=j)
        assert ctx.verify_batch_randomized(0, pks, msgs, known, dst, seed, group=n, rand_bits=bits) == b"\x01"
        assert ctx.verify_batch_randomized(0, pks, msgs, known, dst, bytes(32), group=n, rand_bits=bits) == b"\x00"
        assert ctx.verify_batch_randomized(0, pks, msgs, known, dst, seed, group=n, rand_bits=192 - bits) == b"\x00"
        # the index that enters the scalar is the tuple's index in the CALL, also when the pair sits in a later group
        want = bytearray(b"\x01" * ((n + 15) // 16))
        if i // 16 != j // 16:
            want[i // 16] = want[j // 16] = 0
        assert ctx.verify_batch_randomized(0, pks, msgs, known, dst, seed, group=16, rand_bits=bits) == bytes(want)


def test_randomized_batch_verify_chunked(orc, B, monkeypatch):
    """groups never straddle a staging chunk: BN254_CHUNK=100 with group 48 makes chunks of 96 tuples"""
    monkeypatch.setenv("BN254_CHUNK", "100")
    c = B.Context([0])
    try:
        n = 500
        pks, msgs, sigs, exp = B.scheme.synth_batch(c, n, variant=0, corrupt_every=90)
        seed = b"\x11" * 32
        got = c.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=48)
        assert got == bytes(int(all(exp[g:g + 48])) for g in range(0, n, 48))
        assert got == orc.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=48, nthreads=8)
        got = c.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=300)   # group > chunk
        assert got == b"\x00\x00"
        assert B.scheme.verify_randomized(c, pks, msgs, sigs, group=48) == exp
    finally:
        c.close()


def test_randomized_batch_verify_full_size(ctx, orc, B):
    """2^16 valid tuples in groups of 256: every group accepts; one flipped bit rejects exactly its group.  Then 70001
    tuples (more than one wave of lane pairs: the kernels switch to two tuples per thread sharing one two-pair Miller
    loop, odd tail included) with the same properties, and a sample of groups re-checked against the oracle."""
    big = 70001
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, big, variant=0, n_keys=256, corrupt_every=0)
    seed = b"\xc3" * 32
    n = 1 << 16
    cut = lambda k: (pks[:128 * k], msgs[:k], sigs[:64 * k])
    # aggregate verification at this size also takes two pairs per thread; its un-exponentiated product must equal,
    # byte for byte, the product of three shards small enough for the one-pair-per-thread kernels
    agg = B.scheme.aggregate(ctx, sigs, 0)
    assert ctx.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG) == 1
    whole, bad = ctx.aggregate_partial(0, pks, msgs, agg, B.DST_MIN_SIG)
    cuts = [0, 23001, 47000, big]
    parts = b""
    for k in range(3):
        lo, hi = cuts[k], cuts[k + 1]
        part, b2 = ctx.aggregate_partial(0, pks[128 * lo:128 * hi], msgs[lo:hi], agg if k == 2 else None, B.DST_MIN_SIG)
        parts += part
        bad |= b2
    assert bad == 0 and ctx.fp12_product(parts) == whole
    p, m, s = cut(n)
    assert ctx.verify_batch_randomized(0, p, m, s, B.DST_MIN_SIG, seed, group=256) == b"\x01" * 256
    assert ctx.verify_batch_randomized(0, p, m, s, B.DST_MIN_SIG, seed, group=65536) == b"\x01"
    g = 250
    ng = (big + g - 1) // g
    assert ctx.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=g) == b"\x01" * ng
    assert ctx.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=g, rand_bits=128) == b"\x01" * ng
    for j in (40000, big - 1):
        msgs[j] = msgs[j][:-1] + bytes([msgs[j][-1] ^ 1])
    p, m, s = cut(n)
    got = ctx.verify_batch_randomized(0, p, m, s, B.DST_MIN_SIG, seed, group=256)
    assert got == b"\x01" * 156 + b"\x00" + b"\x01" * 99
    assert ctx.verify_batch_randomized(0, p, m, s, B.DST_MIN_SIG, seed, group=65536) == b"\x00"
    want = bytearray(b"\x01" * ng)
    want[40000 // g] = want[(big - 1) // g] = 0
    got = ctx.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=g)
    assert got == bytes(want)
    # the two-tuples-per-thread path against the oracle: the last 501 tuples as a call of their own would use other
    # scalars (the index enters r_i), so compare through the known-seed forgery instead: it must be accepted at full size
    msgs[40000] = msgs[40000][:-1] + bytes([msgs[40000][-1] ^ 1])
    msgs[big - 1] = msgs[big - 1][:-1] + bytes([msgs[big - 1][-1] ^ 1])
    i, j = 66001, 66002      # items of two different threads (66000/66001 and 66002/66003 share a loop each), same group
    tail = rbv_forgery_at(orc, sigs, seed, i, j)
    got = ctx.verify_batch_randomized(0, pks, msgs, tail, B.DST_MIN_SIG, seed, group=g)
    assert got == b"\x01" * ng
    assert ctx.verify_batch(0, pks[128 * i:128 * (j + 1)], msgs[i:j + 1], tail[64 * i:64 * (j + 1)], B.DST_MIN_SIG) == b"\x00\x00"
    assert ctx.verify_batch_randomized(0, pks, msgs, tail, B.DST_MIN_SIG, bytes(32), group=g)[i // g] == 0


@pytest.fixture(params=[0, 1], ids=["default-shape", "one-thread"])
def big_shape(request, ctx):
    """Big batches run several waves of the lane-pair kernels by default; "lanes" = 1 forces the one-thread kernels
    (k_miller2_fixed, k_miller2, k_miller1, k_miller_twin, k_final_exp) so that they meet the oracle at these sizes too."""
    ctx.set_option("lanes", request.param)
    yield request.param
    ctx.set_option("lanes", 0)


def test_randomized_and_aggregate_beyond_one_lane_wave(ctx, B, big_shape):
    """140 001 tuples: more than the lane-pair kernels hold in one wave; the randomised and aggregate paths run the twin
    kernel (two tuples per thread / lane pair, odd tail) in both shapes.  Valid groups accept, the groups of two
    corrupted tuples reject, and the aggregate product equals, byte for byte, the product of shards that take other kernels."""
    n = 140001
    pks, msgs, sigs, _ = B.scheme.synth_batch(ctx, n, variant=0, n_keys=128, corrupt_every=0)
    seed = b"\x3c" * 32
    g = 500
    ng = (n + g - 1) // g
    agg = B.scheme.aggregate(ctx, sigs, 0)
    whole, bad = ctx.aggregate_partial(0, pks, msgs, agg, B.DST_MIN_SIG)
    parts = b""
    cuts = [0, 30001, 60000, 90001, 120000, n]
    for k in range(5):
        lo, hi = cuts[k], cuts[k + 1]
        part, b2 = ctx.aggregate_partial(0, pks[128 * lo:128 * hi], msgs[lo:hi], agg if k == 0 else None, B.DST_MIN_SIG)
        parts += part
        bad |= b2
    assert bad == 0 and ctx.fp12_product(parts) == whole
    assert ctx.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG) == 1
    assert ctx.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=g) == b"\x01" * ng
    for j in (777, n - 1):
        msgs[j] = msgs[j][:-1] + bytes([msgs[j][-1] ^ 1])
    want = bytearray(b"\x01" * ng)
    want[777 // g] = want[(n - 1) // g] = 0
    assert ctx.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=g) == bytes(want)
    assert ctx.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG) == 0


def rbv_forgery_at(orc, sigs, seed, i, j, bits=64):
    from rbv_util import rbv_forgery
    return rbv_forgery(orc, sigs, seed, bits=bits, i=i, j=j)


def test_verify_full_size_property(ctx, orc, B):
    """BASELINE config 2 size (2^16 tuples, min-sig): accept bits must equal the construction; a 96-tuple
    sample is re-checked against the oracle."""
    n = 1 << 16
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, n, variant=0, n_keys=256)
    got = ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG)
    assert got == exp
    assert sum(exp) == n - n // 64
    lo = 4000
    sl = slice(lo, lo + 96)
    ref = orc.verify_batch(0, pks[lo * 128:(lo + 96) * 128], msgs[sl], sigs[lo * 64:(lo + 96) * 64], B.DST_MIN_SIG, nthreads=8)
    assert got[sl] == ref


def test_hash_full_size_property(ctx, orc):
    """BASELINE config 3 shape at 2^17 messages: a sample against the oracle, and equal messages hash
    equally wherever they sit in the batch (chunk / offset handling)."""
    n = 1 << 17
    rng = random.Random(21)
    base = [rng.randbytes(32) for _ in range(1024)]
    msgs = [base[i % 1024] for i in range(n)]
    out = ctx.hash_to_g1(msgs, QUUX_G1)
    ref = orc.hash_to_g1(base, QUUX_G1, nthreads=8)
    assert out[:64 * 1024] == ref
    assert out[-64 * 1024:] == ref
    mid = (n // 2 // 1024) * 1024
    assert out[64 * mid:64 * (mid + 1024)] == ref


def test_multi_device_ctx(orc, B):
    import ctypes
    nd = B._native.lib().bn254_device_count() if hasattr(B, "_native") else 1
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    c1 = B.Context([0])
    c2 = B.Context([0, 1])
    try:
        pks, msgs, sigs, exp = B.scheme.synth_batch(c1, 700, variant=0, n_keys=16)
        assert c2.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG) == exp
        agg = B.scheme.aggregate(c1, sigs, 0)
        pks2, msgs2, sigs2, _ = B.scheme.synth_batch(c1, 300, variant=0, corrupt_every=0, n_keys=16)
        agg = B.scheme.aggregate(c1, sigs2, 0)
        assert c2.aggregate_verify(0, pks2, msgs2, agg, B.DST_MIN_SIG) == 1
        assert c2.hash_to_g1(msgs, QUUX_G1) == c1.hash_to_g1(msgs, QUUX_G1)
        # randomised batch verification shares out whole groups
        seed = b"\x77" * 32
        want = bytes(int(all(exp[g:g + 32])) for g in range(0, 700, 32))
        assert c2.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=32) == want == c1.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=32)
        # same-key grouping shares out whole keys; the scalars follow the tuple's index in the CALL, whichever device takes it
        counts = [40, 0, 300, 7, 160, 90, 1, 33]
        _, _, kmsgs, kpks, ksigs = _same_key_batch(c1, B, random.Random(5), 0, counts, msg_len=(24,))
        t = sum(counts[:5]) + 17
        ksigs[64 * t:64 * t + 64] = ksigs[64 * (t + 1):64 * (t + 2)]              # key 5 (device 1's share)
        got = c2.verify_same_key_randomized(0, bytes(kpks), counts, kmsgs, bytes(ksigs), B.DST_MIN_SIG, seed)
        assert got == bytes([1, 0, 1, 1, 1, 0, 1, 1]) == c1.verify_same_key_randomized(0, bytes(kpks), counts, kmsgs, bytes(ksigs), B.DST_MIN_SIG, seed)
        assert got == orc.verify_same_key_randomized(0, bytes(kpks), counts, kmsgs, bytes(ksigs), B.DST_MIN_SIG, seed, nthreads=orc.ncpu())
    finally:
        c1.close()
        c2.close()


def test_subgroup_checks(ctx, orc, B):
    """is_torsion_free (g1.rs:852-855, g2.rs:733-736) and the CHECK_SUBGROUP flag of verify_batch."""
    g1, g2 = _points(ctx, orc, 6, 17)
    raw = orc.svdw_g2(b"".join((i + 5).to_bytes(32, "big") + (i + 9).to_bytes(32, "big") for i in range(4)))  # on the curve, outside the subgroup
    ident2 = bytes(64) + bytes(63) + b"\x01"
    from oracle import pymodel as pm
    cof = orc.g2_mul(raw, pm.R_ORDER.to_bytes(32, "big") * 4)          # components in the cofactor subgroup
    mixed = b"".join(orc.g2_sum(g2[128 * i:128 * (i + 1)] + cof[128 * i:128 * (i + 1)]) for i in range(4))
    pts2 = B.G2_GENERATOR + g2 + raw + cof + mixed + ident2 + b"\xff" * 128
    assert ctx.g2_is_torsion_free(pts2) == orc.g2_is_torsion_free(pts2) == b"\x01" * 7 + b"\x00" * 12 + b"\x01\x00"
    assert ctx.g2_is_torsion_free(raw) == b"\x00" * 4 and ctx.g2_is_on_curve(raw) == b"\x01" * 4
    off = bytearray(g1)
    off[40] ^= 1
    pts1 = B.G1_GENERATOR + g1 + bytes(off[:64])
    assert ctx.g1_is_torsion_free(pts1) == orc.g1_is_torsion_free(pts1)
    # a key outside the subgroup is rejected only with the flag
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, 8, variant=0, corrupt_every=0)
    pks = raw[:128] + pks[128:]
    for flags in (0, B.CHECK_ON_CURVE, B.CHECK_SUBGROUP):
        assert ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG, flags=flags) == orc.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG, flags=flags, nthreads=8)
    assert ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG, flags=B.CHECK_SUBGROUP)[1:] == b"\x01" * 7


def test_ragged_sizes_and_long_messages(ctx, orc, B, shape):
    """batch sizes around the CTA / warp boundaries and messages from 0 bytes to 10 KB"""
    rng = random.Random(33)
    for n in (1, 31, 33, 63, 65, 127, 129):
        msgs = [rng.randbytes(rng.choice((0, 1, 31, 32, 55, 56, 63, 64, 65, 119, 120, 200))) for _ in range(n)]
        assert ctx.hash_to_g1(msgs, QUUX_G1) == orc.hash_to_g1(msgs, QUUX_G1, nthreads=8)
    big = [rng.randbytes(10000), b"", rng.randbytes(4097)]
    assert ctx.hash_to_g1(big, QUUX_G1) == orc.hash_to_g1(big, QUUX_G1)
    assert ctx.hash_to_g2(big, QUUX_G2) == orc.hash_to_g2(big, QUUX_G2)
    for n in (1, 33, 65):
        pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, n, variant=0, corrupt_every=4)
        assert ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG) == exp == orc.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG, nthreads=8)
        g1, g2 = _points(ctx, orc, n, n)
        assert ctx.pairing(g1, g2) == orc.pairing_batch(g1, g2, nthreads=8)


def test_profile_and_launch_count(ctx, orc, B):
    """every product call launches kernels of this library (no silent fallback), and the per-launch timer sees them"""
    before = ctx.launch_count
    ctx.profile(True)
    ctx.profile_drain()
    pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, 16, variant=0, corrupt_every=0)
    ctx.profile_drain()
    assert ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG) == exp
    recs = ctx.profile_drain()
    ctx.profile(False)
    names = [r[0] for r in recs]
    # a batch this small runs the latency shape of the two pairing kernels (one warp per pairing, k_w_*) plus the fix-up pass
    # for dead pairs; "wide_max" = 0 sends it to the lane-pair shape
    # (hashing first: in the host pipeline it only waits for the messages, the keys and signatures arrive under it)
    assert names == ["k_hash_g1", "k_g1_decode", "k_g2_decode", "k_w_miller2_fixed", "k_vm_fixup_miller2_fixed", "k_w_final_exp", "k_fp12_is_one"]
    ctx.set_option("wide_max", 0)
    ctx.profile(True)
    try:
        assert ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG) == exp
        names = [r[0] for r in ctx.profile_drain()]
        assert names == ["k_hash_g1", "k_g1_decode", "k_g2_decode", "k_miller2_fixed_l2", "k_final_exp_l2", "k_fp12_is_one"]
    finally:
        ctx.profile(False)
        ctx.set_option("wide_max", 1536)
    assert all(r[2] > 0 for r in recs)
    assert ctx.launch_count > before


def test_aggregate_verify_honours_the_validation_flags(ctx, orc, B):
    """CHECK_SUBGROUP used to be ignored by the aggregate path (and, alone, even switched the on-curve
    test off).  A min-sig key on the twist but outside the order-r subgroup, and an off-curve key, against the oracle."""
    n = 9
    pks, msgs, sigs, _ = B.scheme.synth_batch(ctx, n, variant=0, corrupt_every=0)
    agg = B.scheme.aggregate(ctx, sigs, 0)
    raw = orc.svdw_g2((5).to_bytes(32, "big") + (9).to_bytes(32, "big"))     # on the curve, outside the subgroup
    assert ctx.g2_is_on_curve(raw) == b"\x01" and ctx.g2_is_torsion_free(raw) == b"\x00"
    off = bytearray(pks[:128])
    off[127] ^= 1                                                              # not on the curve
    for bad_pk, verdicts in ((raw, {0: None, 1: None, 2: 0, 3: 0}), (bytes(off), {0: None, 1: 0, 2: 0, 3: 0})):
        Q = pks[:128 * 3] + bad_pk + pks[128 * 4:]
        for flags, want in verdicts.items():
            got = ctx.aggregate_verify(0, Q, msgs, agg, B.DST_MIN_SIG, flags=flags)
            assert got == orc.aggregate_verify(0, Q, msgs, agg, B.DST_MIN_SIG, flags=flags, nthreads=8), flags
            if want is not None:
                assert got == want
    for flags in (0, 1, 2, 3):
        assert ctx.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG, flags=flags) == 1
    # min-pk: the aggregate SIGNATURE is the G2 input
    pks1, msgs1, sigs1, _ = B.scheme.synth_batch(ctx, n, variant=1, corrupt_every=0)
    agg1 = B.scheme.aggregate(ctx, sigs1, 1)
    assert ctx.aggregate_verify(1, pks1, msgs1, agg1, B.DST_MIN_PK, flags=3) == 1
    mixed = orc.g2_sum(agg1 + orc.g2_mul(raw, B.R_ORDER.to_bytes(32, "big")))  # + a point of the cofactor subgroup
    for flags in (0, 1, 2, 3):
        got = ctx.aggregate_verify(1, pks1, msgs1, mixed, B.DST_MIN_PK, flags=flags)
        assert got == orc.aggregate_verify(1, pks1, msgs1, mixed, B.DST_MIN_PK, flags=flags, nthreads=8)
    assert ctx.aggregate_verify(1, pks1, msgs1, mixed, B.DST_MIN_PK, flags=2) == 0
    # the scheme layer validates by default
    assert B.scheme.aggregate_verify(ctx, pks[:128 * 3] + raw + pks[128 * 4:], msgs, agg) == 0
    assert B.scheme.verify(ctx, raw + pks[128:], msgs, sigs)[1:] == b"\x01" * (n - 1)


def test_verify_c5_shape(orc, B):
    """The shapes BASELINE config 5 (2^24 tuples) runs through the double-buffered host pipeline: by default chunks of one wave
    of the lane-pair kernels (66 304 tuples: k_miller2_fixed_l2, k_final_exp_l2, one 896-thread CTA per SM); with "lanes" = 1
    chunks of 2^18 tuples, one thread per pairing (k_miller2_fixed, k_final_exp).  n = 2^18 + 4097: accept bytes equal the construction in both, a 512-tuple slice
    straddling the chunk boundary equals the oracle, and the profiler confirms which kernels launched.  Pageable and pinned
    result buffers take different copy paths."""
    import numpy as np
    import torch
    c = B.Context([0])
    try:
        n = (1 << 18) + 4097
        pks, msgs, sigs, exp = B.scheme.synth_batch(c, n, variant=0, n_keys=64)
        (got), names = launched(c, lambda: c.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG))
        assert got == exp
        assert {"k_miller2_fixed_l2", "k_final_exp_l2", "k_hash_g1"} <= names and "k_miller2_fixed" not in names
        c.set_option("lanes", 1)
        (got), names = launched(c, lambda: c.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG))
        c.set_option("lanes", 0)
        assert got == exp
        assert {"k_miller2_fixed", "k_final_exp", "k_hash_g1"} <= names and not any(x.endswith("_l2") for x in names)
        lo = (1 << 18) - 256
        ref = orc.verify_batch(0, pks[128 * lo:128 * (lo + 512)], msgs[lo:lo + 512], sigs[64 * lo:64 * (lo + 512)], B.DST_MIN_SIG, nthreads=orc.ncpu())
        assert got[lo:lo + 512] == ref
        assert 0 < sum(ref) < 512
        # pinned caller buffers: the zero-staging path of the pipeline, three chunks
        c.set_option("chunk", 100000)
        blob, off = B.pack_msgs(msgs)
        pin = lambda x: torch.frombuffer(bytearray(x), dtype=torch.uint8).pin_memory()
        h_pk, h_sig, h_blob = pin(pks), pin(sigs), pin(blob)
        h_off = torch.tensor(list(off), dtype=torch.int64).pin_memory()
        h_ok = torch.zeros(n, dtype=torch.uint8).pin_memory()
        c.raw("bn254_verify_batch", 0, h_pk.data_ptr(), h_blob.data_ptr(), h_off.data_ptr(), h_sig.data_ptr(), n, B.DST_MIN_SIG, len(B.DST_MIN_SIG), 0,
              h_ok.data_ptr())
        assert bytes(h_ok.numpy()) == exp
        # randomised, same job: groups of 512 over three chunks (twin kernel, one thread per two tuples)
        seed = b"\x5a" * 32
        grp, names = launched(c, lambda: c.verify_batch_randomized(0, pks, msgs, sigs, B.DST_MIN_SIG, seed, group=512))
        assert grp == bytes(int(all(exp[g:g + 512])) for g in range(0, n, 512))
        assert names & {"k_miller_twin", "k_miller_twin_l2"}      # two tuples per thread / lane pair
    finally:
        c.close()


def test_pairing_and_final_exp_beyond_one_wave(ctx, orc, B, big_shape):
    """70 001 pairings: more than one wave of lane pairs, so pairing_batch / miller_loop_batch / final_exponentiation run
    k_miller1 and k_final_exp (one thread per item).  Samples from the head, the middle and the ragged tail against the
    oracle; the profiler confirms which kernels ran."""
    n = 70001
    rng = random.Random(99)
    base = 257
    msgs = [rng.randbytes(16) for _ in range(base)]
    b1, b2 = orc.hash_to_g1(msgs, QUUX_G1, nthreads=8), orc.hash_to_g2(msgs, QUUX_G2, nthreads=8)
    idx = [(i * 31 + i // base) % base for i in range(n)]            # item i pairs P[idx[i]] with Q[(idx[i] * 7 + i // base) % base]
    jdx = [(idx[i] * 7 + i // base) % base for i in range(n)]
    g1 = b"".join(b1[64 * k:64 * k + 64] for k in idx)
    g2 = b"".join(b2[128 * k:128 * k + 128] for k in jdx)
    (gt, st), names = launched(ctx, lambda: ctx.pairing(g1, g2, want_status=True))
    if big_shape == 1:
        assert {"k_miller1", "k_final_exp"} <= names and not any(x.endswith("_l2") for x in names)
    else:
        assert {"k_miller1_l2", "k_final_exp_l2"} <= names and "k_miller1" not in names
    assert st == b"\x01" * n
    for lo in (0, 35000, n - 40):
        sl = slice(lo, lo + 40)
        ref = orc.pairing_batch(g1[64 * lo:64 * (lo + 40)], g2[128 * lo:128 * (lo + 40)], nthreads=8)
        assert gt[384 * lo:384 * (lo + 40)] == ref, lo
    ml = ctx.miller_loop(g1, g2)
    lo = n - 33
    assert ml[384 * lo:] == orc.miller_loop_batch(g1[64 * lo:], g2[128 * lo:], nthreads=8)
    fe, names = launched(ctx, lambda: ctx.final_exponentiation(ml))
    assert ("k_final_exp" if big_shape == 1 else "k_final_exp_l2") in names
    assert fe == gt
    # min-pk verification in one chunk of this size: k_miller2 / k_miller2_l2 (both pairs variable)
    m = 66400
    c = B.Context([0])
    try:
        c.set_option("lanes", big_shape)
        pks, vm, sigs, exp = B.scheme.synth_batch(c, m, variant=1, n_keys=32)
        got, names = launched(c, lambda: c.verify_batch(1, pks, vm, sigs, B.DST_MIN_PK))
        assert got == exp and ("k_miller2" if big_shape == 1 else "k_miller2_l2") in names and "k_hash_g2" in names
        lo = m - 48
        assert got[lo:] == orc.verify_batch(1, pks[64 * lo:], vm[lo:], sigs[128 * lo:], B.DST_MIN_PK, nthreads=8)
    finally:
        c.close()


def test_aggregate_exchange_single_rank(ctx, orc, B):
    """bn254_aggregate_verify_dist on a lone rank (no communicator needed): product tree -> send record -> post-gather
    kernel, the same launches a rank of an 8-GPU job makes.  Verdicts against the oracle."""
    n = 300
    pks, msgs, sigs, _ = B.scheme.synth_batch(ctx, n, variant=0, corrupt_every=0, n_keys=16)
    agg = B.scheme.aggregate(ctx, sigs, 0)
    ok, names = launched(ctx, lambda: ctx.aggregate_verify_dist(0, pks, msgs, agg, B.DST_MIN_SIG))
    assert ok == 1 == orc.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG, nthreads=8)
    assert {"k_fp12_reduce", "k_agg_rec_flag", "k_w_gather_finish"} <= names      # the lone final exponentiation: the latency shape
    ctx.set_option("wide_max", 0)
    try:
        ok, names = launched(ctx, lambda: ctx.aggregate_verify_dist(0, pks, msgs, agg, B.DST_MIN_SIG))
        assert ok == 1 and "k_gather_finish_l2" in names
    finally:
        ctx.set_option("wide_max", 1536)
    msgs[7] = b"tampered"
    assert ctx.aggregate_verify_dist(0, pks, msgs, agg, B.DST_MIN_SIG) == 0
    assert ctx.aggregate_verify_dist(0, b"\xff" * 128 + pks[128:], msgs, agg, B.DST_MIN_SIG) == 0
    with pytest.raises(B.Bn254Error):
        ctx.aggregate_verify_dist(0, pks, msgs, None, B.DST_MIN_SIG)      # rank 0 must bring the aggregate signature


def test_nccl_exchange_in_multi_device_ctx(orc, B):
    """A ctx over two devices: shards on both, ncclAllGather inside the library, one post-gather kernel on device 0."""
    nd = B._native.lib().bn254_device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    c1, c2 = B.Context([0]), B.Context([0, 1])
    try:
        assert c2.comm_nranks == 2
        n = 501
        pks, msgs, sigs, _ = B.scheme.synth_batch(c1, n, variant=0, corrupt_every=0, n_keys=16)
        agg = B.scheme.aggregate(c1, sigs, 0)
        assert c2.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG) == 1
        assert c2.aggregate_partial(0, pks, msgs, agg, B.DST_MIN_SIG) == c1.aggregate_partial(0, pks, msgs, agg, B.DST_MIN_SIG)
        msgs[400] = b"tampered"                                          # device 1's shard
        assert c2.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG) == 0 == orc.aggregate_verify(0, pks, msgs, agg, B.DST_MIN_SIG, nthreads=8)
    finally:
        c1.close()
        c2.close()


def test_verify_with_prepared_keys(ctx, orc, B, shape):
    """bn254_g2_prepare / bn254_verify_batch_prepared (G2Prepared, pairings.rs:609-660): the exported line tables equal the
    oracle's, and verification against prepared keys gives the accept bytes of bn254_verify_batch on the expanded key array --
    which equal the oracle's -- with unusable keys (identity, undecodable, outside the subgroup when asked) rejecting their tuples."""
    nk, n = 9, 230
    sks = [B.scheme.SecretKey(B.scheme.scalar_from_seed(77, i)) for i in range(nk)]
    keys = bytearray(B.scheme.public_keys(ctx, sks, 0))
    ident2 = bytes(64) + bytes(63) + b"\x01"
    raw = orc.svdw_g2((5).to_bytes(32, "big") + (9).to_bytes(32, "big"))        # on the twist, outside the subgroup
    keys[128 * 6:128 * 7] = ident2
    keys[128 * 7:128 * 8] = b"\xff" * 128
    keys[128 * 8:128 * 9] = raw
    keys = bytes(keys)
    rng = random.Random(41)
    kidx = [rng.randrange(nk) for _ in range(n)]
    msgs = [rng.randbytes(rng.choice((0, 5, 32, 77))) for _ in range(n)]
    sc = b"".join(sks[k].to_bytes() for k in kidx)
    sigs = bytearray(ctx.g1_mul(ctx.hash_to_g1(msgs, B.DST_MIN_SIG), sc))
    for j in range(5, n, 17):                                                   # a few wrong signatures
        sigs[64 * j:64 * (j + 1)] = sigs[64 * (j - 1):64 * j]
    sigs = bytes(sigs)
    pks = b"".join(keys[128 * k:128 * (k + 1)] for k in kidx)
    for flags in (0, B.CHECK_SUBGROUP):
        prep = ctx.g2_prepare(keys, flags)
        assert len(prep) == nk
        tabs, st = prep.export()
        ref_tabs, ref_st = orc.g2_prepare(keys)
        want_st = bytearray(ref_st)
        if flags:
            want_st[8] = 0
        assert st == bytes(want_st)
        for k in range(nk):
            if st[k]:
                assert tabs[88 * 192 * k:88 * 192 * (k + 1)] == ref_tabs[88 * 192 * k:88 * 192 * (k + 1)], k
        got, names = launched(ctx, lambda: ctx.verify_batch_prepared(prep, kidx, msgs, sigs, B.DST_MIN_SIG))
        assert got == ctx.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG, flags=flags)
        assert got == orc.verify_batch(0, pks, msgs, sigs, B.DST_MIN_SIG, flags=flags, nthreads=8)
        if shape == 4:      # the latency shape: the keys of the tuples are fetched and the ordinary wide loop runs
            assert {"k_g2_fetch", "k_w_miller2_fixed"} <= names
        else:
            assert names & {"k_miller2_prep", "k_miller2_prep_l2"}
        assert 0 < sum(got) < n and all(got[i] == 0 for i in range(n) if kidx[i] in (6, 7) or (flags and kidx[i] == 8))
        prep.close()
    prep = ctx.g2_prepare(keys[:128 * 6])
    with pytest.raises(B.Bn254Error):
        ctx.verify_batch_prepared(prep, [6] + kidx[1:], msgs, sigs, B.DST_MIN_SIG)      # key index out of range
    assert ctx.verify_batch_prepared(prep, [], [], b"", B.DST_MIN_SIG) == b""
    prep.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_fast_aggregate_verify_batch(ctx, orc, B, variant, shape):
    """bn254_fast_aggregate_verify_batch (FastAggregateVerify of the IETF draft: many signers, one message per group): verdicts
    equal the oracle's -- valid committees accept; an empty key set, an identity key, an undecodable key, a missing signer, a
    tampered message and two keys that cancel reject."""
    rng = random.Random(61 + variant)
    dst = B.DST_MIN_SIG if variant == 0 else B.DST_MIN_PK
    rpk, rsig = (128, 64) if variant == 0 else (64, 128)
    sizes = [1, 3, 17, 0, 5, 4, 64, 2, 6, 2]
    G = len(sizes)
    msgs = [rng.randbytes(rng.choice((0, 7, 32, 100))) for _ in range(G)]
    R = B.R_ORDER
    keys, sigs = b"", b""
    ident_pk = (bytes(64) + bytes(63) + b"\x01") if variant == 0 else ((0).to_bytes(32, "big") + (1).to_bytes(32, "big"))
    ident_sig = ((0).to_bytes(32, "big") + (1).to_bytes(32, "big")) if variant == 0 else (bytes(64) + bytes(63) + b"\x01")
    for g, k in enumerate(sizes):
        sks = [B.scheme.SecretKey(rng.randrange(1, R)) for _ in range(k)]
        if g == 9:      # two keys that cancel: the aggregate key is the identity
            sks[1] = B.scheme.SecretKey(R - sks[0].value)
        pk = bytearray(B.scheme.public_keys(ctx, sks, variant)) if k else bytearray()
        if k:
            per = B.scheme.sign(ctx, sks, [msgs[g]] * k, variant)
            agg = B.scheme.aggregate(ctx, per, variant)
        else:
            agg = ident_sig
        if g == 4:
            pk[rpk * 2:rpk * 3] = ident_pk          # an identity key inside the committee
        if g == 5:
            pk[rpk:2 * rpk] = b"\xff" * rpk          # an undecodable key
        if g == 7:      # one signer missing from the aggregate
            agg = B.scheme.sign(ctx, sks[:1], [msgs[g]], variant)
        keys += bytes(pk)
        sigs += agg
    msgs[8] = msgs[8] + b"!"                         # tampered message
    for flags in (0, 3):
        got = ctx.fast_aggregate_verify_batch(variant, keys, sizes, msgs, sigs, dst, flags=flags)
        assert got == orc.fast_aggregate_verify_batch(variant, keys, sizes, msgs, sigs, dst, flags=flags), flags
        assert got == bytes([1, 1, 1, 0, 0, 0, 1, 0, 0, 0])
    assert ctx.fast_aggregate_verify_batch(variant, b"", [], [], b"", dst) == b""


def _same_key_batch(ctx, B, rng, variant, counts, msg_len=(0, 7, 32, 100)):
    """keys, messages and signatures of a same-key batch: key k signs the next counts[k] messages"""
    R = B.R_ORDER
    nk = len(counts)
    sks = [B.scheme.SecretKey(rng.randrange(1, R)) for _ in range(nk)]
    key_of = [k for k, c in enumerate(counts) for _ in range(c)]
    msgs = [i.to_bytes(4, "little") + rng.randbytes(rng.choice(msg_len)) for i in range(len(key_of))]
    pks = B.scheme.public_keys(ctx, sks, variant)
    sigs = B.scheme.sign(ctx, [sks[k] for k in key_of], msgs, variant) if key_of else b""
    return sks, key_of, msgs, bytearray(pks), bytearray(sigs)


@pytest.fixture(params=[0, 1], ids=["scaled-points", "bucket-msm"])
def sk_sums(request, ctx):
    """Both ways the same-key path forms a key's two sums: every point scaled, then segmented sums (sk_msm_min = 0: never the
    bucket path) / bucket-method multi-scalar multiplications (sk_msm_min = 1: always)."""
    ctx.set_option("sk_msm_min", request.param)
    yield request.param
    ctx.set_option("sk_msm_min", 128)


@pytest.mark.parametrize("variant", [0, 1])
def test_same_key_randomized_vs_oracle(ctx, orc, B, variant, shape, sk_sums):
    """bn254_verify_same_key_randomized (many messages under few keys: ONE Miller loop and ONE final exponentiation per key over
    the randomised sums of that key's signatures and message points): verdicts equal the oracle's -- clean keys accept; a key
    without tuples, a tampered message, a signature under another key, an identity signature, an identity key and an
    undecodable key reject exactly their key -- with both scalar widths and with validation on."""
    rng = random.Random(71 + variant)
    dst = B.DST_MIN_SIG if variant == 0 else B.DST_MIN_PK
    rpk, rsig = (128, 64) if variant == 0 else (64, 128)
    counts = [1, 5, 0, 37, 4, 2, 16, 3, 3, 9, 1]
    sks, key_of, msgs, pks, sigs = _same_key_batch(ctx, B, rng, variant, counts)
    start = [sum(counts[:k]) for k in range(len(counts))]
    ident_pk = (bytes(64) + bytes(63) + b"\x01") if variant == 0 else ((0).to_bytes(32, "big") + (1).to_bytes(32, "big"))
    ident_sig = ((0).to_bytes(32, "big") + (1).to_bytes(32, "big")) if variant == 0 else (bytes(64) + bytes(63) + b"\x01")
    seed = rng.randbytes(32)
    clean = bytes(int(c > 0) for c in counts)
    assert ctx.verify_same_key_randomized(variant, bytes(pks), counts, msgs, bytes(sigs), dst, seed) == clean
    msgs[start[3] + 20] += b"?"                                                         # key 3: tampered message
    t4 = start[4] + 1                                                                   # key 4: a signature made by key 5
    sigs[t4 * rsig:(t4 + 1) * rsig] = B.scheme.sign(ctx, [sks[5]], [msgs[t4]], variant)
    t6 = start[6] + 15                                                                  # key 6: identity signature (last tuple)
    sigs[t6 * rsig:(t6 + 1) * rsig] = ident_sig
    pks[7 * rpk:8 * rpk] = ident_pk                                                     # key 7: identity key
    pks[8 * rpk:9 * rpk] = b"\xff" * rpk                                                # key 8: undecodable key
    want = bytes([1, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1])
    for flags, bits in ((0, 64), (3, 64), (1, 128)):
        got = ctx.verify_same_key_randomized(variant, bytes(pks), counts, msgs, bytes(sigs), dst, seed, rand_bits=bits, flags=flags)
        ref = orc.verify_same_key_randomized(variant, bytes(pks), counts, msgs, bytes(sigs), dst, seed, rand_bits=bits, flags=flags, nthreads=orc.ncpu())
        assert got == ref == want, (flags, bits)
    # the scheme layer turns per-key verdicts into per-tuple ones: rejected keys are re-verified exactly
    per_tuple = B.scheme.verify_same_key(ctx, bytes(pks), counts, msgs, bytes(sigs), variant, flags=0, seed=seed)
    key_of = [k for k, c in enumerate(counts) for _ in range(c)]
    exact = ctx.verify_batch(variant, b"".join(bytes(pks[k * rpk:(k + 1) * rpk]) for k in key_of), msgs, bytes(sigs), dst)
    assert per_tuple == exact and 0 < sum(exact) < len(exact)
    assert ctx.verify_same_key_randomized(variant, b"", [], [], b"", dst, seed) == b""
    assert ctx.verify_same_key_randomized(variant, bytes(pks[:2 * rpk]), [0, 0], [], b"", dst, seed) == b"\x00\x00"
    with pytest.raises(B.Bn254Error):
        ctx.verify_same_key_randomized(variant, bytes(pks), counts[:-1], msgs, bytes(sigs), dst, seed)


def test_same_key_randomized_multi_chunk_and_predicted_scalars(ctx, orc, B):
    """The host pipeline of the same-key path: with a small chunk the keys stream through several staging rounds (a key is never
    split: the 600-tuple key exceeds the chunk and goes alone) and the verdicts equal the one-chunk call and the construction.
    Soundness depends on the seed being unpredictable: a forger who KNOWS the seed cancels two errors (sig_a + D, sig_b - (r_a /
    r_b) D) and is accepted under that seed, rejected under any other -- the oracle agrees on both."""
    from rbv_util import _rbv_scalar
    rng = random.Random(83)
    counts = [600, 3, 0, 250, 40, 7, 130, 1]
    sks, key_of, msgs, pks, sigs = _same_key_batch(ctx, B, rng, 0, counts, msg_len=(28,))
    start = [sum(counts[:k]) for k in range(len(counts))]
    bad = start[3] + 101
    sigs[64 * bad:64 * bad + 64] = sigs[64 * (bad + 1):64 * (bad + 2)]
    want = bytes([1, 1, 0, 0, 1, 1, 1, 1])
    seed = rng.randbytes(32)
    one = ctx.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(sigs), B.DST_MIN_SIG, seed)
    assert one == want
    c = B.Context([0])
    try:
        c.set_option("chunk", 256)
        c.set_option("sk_msm_min", 0)
        many, names = launched(c, lambda: c.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(sigs), B.DST_MIN_SIG, seed))
        # without the profiler the per-key tail of a chunk runs on its own stream beside the next chunk
        overlapped = [c.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(sigs), B.DST_MIN_SIG, seed) for _ in range(3)]
    finally:
        c.close()
    assert overlapped == [want] * 3
    assert many == want
    assert {"k_sk_sum_g1p", "k_sk_finish_g1", "k_sk_key_flags", "k_rbv_scale_g1p"} <= names, names
    # the same job with the sums as bucket-method multi-scalar multiplications (default for segments this long), 128-bit scalars too
    msm, names = launched(ctx, lambda: ctx.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(sigs), B.DST_MIN_SIG, seed))
    assert msm == want and {"k_sk_digits", "k_sk_bucket_g1", "k_sk_window_g1", "k_sk_combine_g1"} <= names, names
    for bits in (64, 128):
        got = ctx.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(sigs), B.DST_MIN_SIG, seed[::-1], rand_bits=bits, flags=3)
        assert got == want == orc.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(sigs), B.DST_MIN_SIG, seed[::-1], rand_bits=bits, flags=3,
                                                             nthreads=orc.ncpu())
    # predicted scalars: key 4's tuples a, b get sig_a + D and sig_b - [r_a / r_b] D
    R = B.R_ORDER
    a, b = start[4] + 2, start[4] + 9
    ra, rb = _rbv_scalar(seed, a, 64), _rbv_scalar(seed, b, 64)
    D = B.scheme.sign(ctx, [B.scheme.SecretKey(12345)], [b"delta"], 0)
    c = (R - ra * pow(rb, -1, R)) % R
    cD = ctx.g1_mul(D, c.to_bytes(32, "big"))
    forged = bytearray(sigs)
    forged[64 * bad:64 * bad + 64] = B.scheme.sign(ctx, [sks[3]], [msgs[bad]], 0)        # repair key 3
    forged[64 * a:64 * a + 64] = ctx.g1_sum(bytes(forged[64 * a:64 * a + 64]) + D)
    forged[64 * b:64 * b + 64] = ctx.g1_sum(bytes(forged[64 * b:64 * b + 64]) + cD)
    got = ctx.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(forged), B.DST_MIN_SIG, seed)
    assert got == bytes([1, 1, 0, 1, 1, 1, 1, 1])                                        # accepted under the KNOWN seed ...
    assert got == orc.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(forged), B.DST_MIN_SIG, seed, nthreads=orc.ncpu())
    other = ctx.verify_same_key_randomized(0, bytes(pks), counts, msgs, bytes(forged), B.DST_MIN_SIG, bytes(32))
    assert other == bytes([1, 1, 0, 1, 0, 1, 1, 1])                                      # ... and caught under any other
    exact = ctx.verify_batch(0, b"".join(bytes(pks[128 * k:128 * k + 128]) for k in key_of), msgs, bytes(forged), B.DST_MIN_SIG)
    assert exact[a] == 0 and exact[b] == 0 and sum(exact) == len(key_of) - 2


def test_wide_shape_boundary(ctx, orc, B):
    """Shape 0 picks the wide (one warp per pairing) kernels for launches of at most "wide_max" pairings and the lane-pair
    kernels above: both sides of the boundary against the oracle, min-sig and min-pk, standalone pairings and a lone final
    exponentiation included, and the profiler confirms which kernels ran."""
    ctx.set_option("wide_max", 40)
    try:
        for variant, dst in ((0, B.DST_MIN_SIG), (1, B.DST_MIN_PK)):
            for n, wide in ((40, True), (41, False)):
                pks, msgs, sigs, exp = B.scheme.synth_batch(ctx, n, variant=variant, corrupt_every=7)
                got, names = launched(ctx, lambda: ctx.verify_batch(variant, pks, msgs, sigs, dst))
                assert got == exp == orc.verify_batch(variant, pks, msgs, sigs, dst, nthreads=8)
                assert any(x.startswith("k_w_") for x in names) == wide, (variant, n, names)
                assert any(x.endswith("_l2") for x in names) == (not wide), (variant, n, names)
        g1, g2 = _points(ctx, orc, 41, 41)
        gt_w, names = launched(ctx, lambda: ctx.pairing(g1[:64 * 40], g2[:128 * 40]))
        assert {"k_w_miller1", "k_w_final_exp"} <= names
        gt_l, names = launched(ctx, lambda: ctx.pairing(g1, g2))
        assert {"k_miller1_l2", "k_final_exp_l2"} <= names
        ref = orc.pairing_batch(g1, g2, nthreads=8)
        assert gt_l == ref and gt_w == ref[:384 * 40]
        ml = orc.miller_loop_batch(g1[:64], g2[:128])
        fe, names = launched(ctx, lambda: ctx.final_exponentiation(ml))
        assert fe == orc.final_exponentiation(ml) and "k_w_final_exp" in names
    finally:
        ctx.set_option("wide_max", 1536)
