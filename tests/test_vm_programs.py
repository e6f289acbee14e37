"""The pairing slot machine (bls_bn254_b200/csrc/vm.cuh + the generated vm_prog.cuh) before a GPU sees it:
  1. every routine of tools/gen_vm_programs.py runs in the generator's big-integer simulator against oracle/pymodel.py;
  2. the committed vm_prog.cuh is what the generator emits (no drift);
  3. the C++ interpreter itself, compiled for the host with two emulated lanes per pairing (tests/hostsim), reproduces the C
     oracle byte for byte: Miller loops of all three families, final exponentiation, whole pairings."""
import ctypes as C
import pathlib
import random
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
DST = b"QUUX-V01-CS02-with-BN254G1_XMD:SHA-256_SVDW_RO_"


@pytest.fixture(scope="module")
def gen():
    import gen_vm_programs as g
    consts, cidx = g.build_constants()
    return g, consts, g.routines(cidx)


def _flat(f):
    return [f[0][0], f[0][1], f[0][2], f[1][0], f[1][1], f[1][2]]


def test_programs_in_the_simulator(gen):
    from oracle import pymodel as m
    g, consts, R = gen
    rng = random.Random(5)
    for trial in range(2):
        P0, P1 = (m.g1_mul(m.G1_GEN, rng.randrange(1, m.R_ORDER)) for _ in range(2))
        Q0, Q1 = (m.g2_mul(m.G2_GEN, rng.randrange(1, m.R_ORDER)) for _ in range(2))
        for fam, pairs in (("M2F", [(P0, m.g2_neg(m.G2_GEN)), (P1, Q1)]), ("M1", [(P1, Q1)]), ("M2V", [(P0, Q0), (P1, Q1)])):
            sim = g.Sim(32, consts)
            sim.scalars = [P0[0], P0[1], P1[0], P1[1]]
            sim.inputs = [Q0[0], Q0[1], Q1[0], Q1[1]]
            assert g.sim_miller(sim, R, fam) == _flat(m.multi_miller_loop(pairs)), fam
        f = m.multi_miller_loop([(P1, Q1)])
        sim = g.Sim(32, consts)
        for i, c in enumerate(_flat(f)):
            sim.s[i] = c
        assert g.sim_final_exp(sim, R) == _flat(m.final_exponentiation(f))
    # the known-answer test of the reference (pairings.rs:970-980) through the simulator
    sim = g.Sim(32, consts)
    sim.scalars = [0, 0, m.G1_GEN[0], m.G1_GEN[1]]
    sim.inputs = [(0, 0), (0, 0), m.G2_GEN[0], m.G2_GEN[1]]
    g.sim_miller(sim, R, "M1")
    assert g.sim_final_exp(sim, R) == _flat(m.pairing(m.G1_GEN, m.G2_GEN))


def test_wide_schedules_in_the_simulator(gen):
    """the wide machine's programs: every routine after move elimination, post splitting and list scheduling, executed bundle by
    bundle (all units read, then all units write) on big integers, against oracle/pymodel.py"""
    from oracle import pymodel as m
    g, consts, _ = gen
    _, cidx = g.build_constants()
    R = g.routines_wide(cidx)
    for a in R.values():       # a bundle never holds more instructions than the warp has units, and every instruction is placed once
        assert all(1 <= len(b) <= g.W_UNITS for b in a.bundles)
        assert sorted(i for b in a.bundles for i in b) == list(range(len(a.ins)))
    rng = random.Random(6)
    P0, P1 = (m.g1_mul(m.G1_GEN, rng.randrange(1, m.R_ORDER)) for _ in range(2))
    Q0, Q1 = (m.g2_mul(m.G2_GEN, rng.randrange(1, m.R_ORDER)) for _ in range(2))
    for fam, pairs in (("M2F", [(P0, m.g2_neg(m.G2_GEN)), (P1, Q1)]), ("M1", [(P1, Q1)]), ("M2V", [(P0, Q0), (P1, Q1)])):
        sim = g.WSim(consts)
        sim.set_item([P0[0], P0[1], P1[0], P1[1]], [Q0[0], Q0[1], Q1[0], Q1[1]])
        assert g.sim_miller(sim, R, fam) == _flat(m.multi_miller_loop(pairs)), fam
    f = m.multi_miller_loop([(P1, Q1)])
    sim = g.WSim(consts)
    for i, c in enumerate(_flat(f)):
        sim.s[i] = c
    assert g.sim_final_exp(sim, R) == _flat(m.final_exponentiation(f))


def test_committed_programs_match_the_generator(gen, tmp_path):
    g = gen[0]
    for name, emit in (("vm_prog.cuh", g.emit), ("vmw_prog.cuh", g.emit_wide)):
        path = ROOT / "bls_bn254_b200" / "csrc" / name
        before = path.read_text()
        emit()
        after = path.read_text()
        assert before == after, "bls_bn254_b200/csrc/%s is stale: run python tools/gen_vm_programs.py" % name


def test_window_digits_are_the_kernels(gen):
    """the C++ driver reads k::X_WNAF4 / k::ATE_NAF; the simulator's digit strings must denote the same integers"""
    g = gen[0]
    x = 0
    for d in g.wnaf4(g.X_PARAM):
        x = 2 * x + d
    assert x == g.X_PARAM
    assert sum(d << i for i, d in enumerate(g.ATE_NAF)) == 6 * g.X_PARAM + 2 and len(g.ATE_NAF) == 66


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "tests" / "hostsim")])
    lib = C.CDLL(str(ROOT / "tests" / "hostsim" / "libhostsim.so"))
    lib.sim_vm.argtypes = [C.c_int, C.c_int] + [C.c_char_p] * 6
    lib.sim_vm_wide.argtypes = [C.c_int, C.c_int] + [C.c_char_p] * 6
    return lib


def _neg_g2(orc):
    P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    g = orc.g2_generator()
    neg = lambda b: ((P - int.from_bytes(b, "big")) % P).to_bytes(32, "big")
    return g[:64] + neg(g[64:96]) + neg(g[96:])


@pytest.mark.parametrize("machine", ["sim_vm", "sim_vm_wide"])
def test_interpreter_on_emulated_lanes_vs_oracle(sim, orc, kats, machine):
    """sim_vm: the slot machine on two emulated lanes; sim_vm_wide: the wide machine on 32 emulated lanes (16 units, bundles)"""
    run = getattr(sim, machine)
    rng = random.Random(12)
    msgs = [rng.randbytes(20) for _ in range(4)]
    g1 = orc.hash_to_g1(msgs, DST)
    g2 = orc.hash_to_g2(msgs, DST)
    out = C.create_string_buffer(384)
    p0, p1, q0, q1 = g1[:64], g1[64:128], g2[:128], g2[128:256]
    # one pair, Miller loop only / whole pairing
    run(0, 0, None, None, p1, q1, None, out)
    assert out.raw == orc.miller_loop_batch(p1, q1)
    run(0, 1, None, None, p1, q1, None, out)
    assert out.raw == orc.pairing_batch(p1, q1)
    # two pairs with the fixed -g2 table (the verification shape)
    run(1, 0, p0, None, p1, q1, None, out)
    assert out.raw == orc.multi_miller_loop(p0 + p1, _neg_g2(orc) + q1)
    # two variable pairs
    run(2, 0, p0, q0, p1, q1, None, out)
    assert out.raw == orc.multi_miller_loop(p0 + p1, q0 + q1)
    # final exponentiation alone
    ml = orc.miller_loop_batch(p0, q0)
    run(-1, 1, None, None, None, None, ml, out)
    assert out.raw == orc.final_exponentiation(ml)
    # pairing(g1, g2) == Gt::generator() (pairings.rs:970-980)
    run(0, 1, None, None, orc.g1_generator(), orc.g2_generator(), None, out)
    assert out.raw == b"".join(bytes.fromhex(x) for x in kats["gt_generator"])


def test_wide_lazy_linear_combination(sim):
    """vm.cuh: wacc_mad / wacc_reduce -- the unreduced signed accumulator of the wide machine's gathers and its single reduction
    (quotient estimate from the top 64 bits) against big integers, at the extremes of the bound (12 terms, |coefficient| <= 40,
    values 0, 1, p - 1) and at random"""
    P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    sim.sim_wide_lincomb.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_char_p]
    rng = random.Random(31)
    out = C.create_string_buffer(32)

    def check(vals, coefs):
        buf = b"".join(v.to_bytes(32, "big") for v in vals)
        arr = (C.c_int * len(coefs))(*coefs)
        sim.sim_wide_lincomb(buf, arr, len(coefs), out)
        assert int.from_bytes(out.raw, "big") == sum(c * v for c, v in zip(coefs, vals)) % P, (vals, coefs)

    check([], [])
    for v in (0, 1, P - 1):
        for c in (40, -40, 1, -1):
            check([v] * 12, [c] * 12)
            check([v], [c])
    check([P - 1] * 12, [40, -40] * 6)
    check([P - 1] * 6 + [0] * 6, [40] * 6 + [-40] * 6)
    for _ in range(3000):
        n = rng.randrange(1, 13)
        vals = [rng.choice((rng.randrange(P), P - 1 - rng.randrange(4), rng.randrange(4))) for _ in range(n)]
        coefs = [rng.choice((1, -1, 2, -2, 3, 4, 9, -9, 18, 27, 36, -36, 40, -40)) for _ in range(n)]
        check(vals, coefs)
    # values whose combination lands next to a multiple of p (quotient estimate one short -> the conditional subtractions)
    for k in range(1, 400, 7):
        check([P - 1, 1], [k % 41 or 1, (k * 3) % 41 or 1])
