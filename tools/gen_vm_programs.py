#!/usr/bin/env python3
"""Assembler + program generator for the pairing "slot machine" (bls_bn254_b200/csrc/vm.cuh).

The pairing kernels of round 1 were call sequences of ~350-instruction Fp2 leaves whose operands lived in per-thread LOCAL
memory (1.6-2.7 KB per thread, L1 hit rate 50-65 %, 29 GB of DRAM write-back per launch).  The slot machine keeps the same
leaves but moves every operand into SHARED memory and replaces the tower code by data:

  * a lane pair owns one pairing; a SLOT is one Fp2 value, 32 bytes per lane, addressed [slot][half][lane] so that every
    LDS.128 / STS.128 of a warp is conflict free;
  * an INSTRUCTION is 16 bytes in constant memory: an opcode, a destination slot and up to twelve operand descriptors.  One
    instruction is a whole fused step of the tower -- R = 2^k * ( xi^e * ((sum X_i) * (sum Y_j) + sum Q_k) + sum R_l ) --
    so the Karatsuba glue (pre-additions, post-subtractions, multiplications by xi) rides in the shadow of the multiplier
    instead of being separate load / add / store round trips;
  * a ROUTINE (one Miller doubling iteration, one cyclotomic squaring, ...) is a straight-line list of instructions; loops
    over NAF digits / window digits stay in the kernel's C++ driver.

This script is the only place that knows the tower algebra of the slot machine: it emits csrc/vm_prog.cuh (programs,
constant table, routine directory, slot count).  tests/test_vm_programs.py runs every routine in the Python simulator
below against oracle/pymodel.py (big integers), and the host build of vm.cuh against the C oracle, before a GPU sees it.

  python tools/gen_vm_programs.py            # (re)write bls_bn254_b200/csrc/vm_prog.cuh
"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R256 = 1 << 256

# ---------------------------------------------------------------------------------------------- ISA
OP_MUL, OP_SQR, OP_MULFP, OP_MULC, OP_MULFPC, OP_LIN, OP_LDC, OP_LDG, OP_INV, OP_FILL, OP_SPILL = range(1, 12)
OPNAMES = {OP_MUL: "MUL", OP_SQR: "SQR", OP_MULFP: "MULFP", OP_MULC: "MULC", OP_MULFPC: "MULFPC", OP_LIN: "LIN", OP_LDC: "LDC",
           OP_LDG: "LDG", OP_INV: "INV", OP_FILL: "FILL", OP_SPILL: "SPILL"}
NEG, XI = 0x80, 0x40

# item scalars (Fp, shared by both lanes) and item Fp2 inputs, by index
SC_P0X, SC_P0Y, SC_P1X, SC_P1Y = 0, 1, 2, 3
IN_Q0X, IN_Q0Y, IN_Q1X, IN_Q1Y = 0, 1, 2, 3


class T:
    """operand term: slot with optional negation / multiplication by xi = 9 + u (and a small coefficient for LIN)"""

    def __init__(self, slot, neg=False, xi=False, coef=1, rell=False, relv=False):
        self.slot, self.neg, self.xi, self.coef = slot, neg, xi, coef
        self.rell, self.relv = rell, relv       # wide machine only: slot += 3 * line / slot += 6 * var at run time

    def __neg__(self):
        return T(self.slot, not self.neg, self.xi, self.coef, self.rell, self.relv)

    def x(self):
        return T(self.slot, self.neg, True, self.coef, self.rell, self.relv)

    def __mul__(self, k):
        return T(self.slot, self.neg != (k < 0), self.xi, self.coef * abs(k), self.rell, self.relv)

    __rmul__ = __mul__

    def byte(self):
        assert 0 <= self.slot < 64 and not (self.rell or self.relv)
        return self.slot | (NEG if self.neg else 0) | (XI if self.xi else 0)

    def half(self):
        """16-bit descriptor of the wide machine: slot(10) | NEG | XI | RELL | RELV | coef - 1 (2)"""
        assert 0 <= self.slot < 1024 and 1 <= self.coef <= 4
        return (self.slot | int(self.neg) << 10 | int(self.xi) << 11 | int(self.rell) << 12 | int(self.relv) << 13 | (self.coef - 1) << 14)


def terms(x):
    if isinstance(x, T):
        return [x]
    if isinstance(x, int):
        return [T(x)]
    return [t if isinstance(t, T) else T(t) for t in x]


class Ins:
    def __init__(self, op, d=0, xs=(), ys=(), q=(), r=(), xi=False, dbl=0, conjx=False, rel=False, aux=0, cidx=None):
        self.op, self.d, self.xs, self.ys, self.q, self.r = op, d, terms(xs), terms(ys), terms(q), terms(r)
        self.xi, self.dbl, self.conjx, self.rel, self.aux, self.cidx = xi, dbl, conjx, rel, aux, cidx

    def encode(self):
        nx, ny, nq, nr = len(self.xs), len(self.ys), len(self.q), len(self.r)
        assert nx <= 7 and ny <= 7 and nq <= 3 and nr <= 7 and 0 <= self.dbl <= 3 and 0 <= self.aux < 64 and 0 <= self.d < 64
        w0 = (self.op | self.d << 4 | nx << 10 | ny << 13 | nq << 16 | nr << 18 | int(self.xi) << 21 | self.dbl << 22 |
              int(self.conjx) << 24 | int(self.rel) << 25 | self.aux << 26)
        if self.op == OP_LIN:
            assert nx <= 8 and not (ny or nq or nr)
            desc = [t.byte() for t in self.xs] + [0] * (8 - nx)
            coefs = 0
            for i, t in enumerate(self.xs):
                assert 1 <= t.coef <= 4
                coefs |= (t.coef - 1) << (2 * i)
            words = [w0, int.from_bytes(bytes(desc[0:4]), "little"), int.from_bytes(bytes(desc[4:8]), "little"), coefs]
            return words
        assert all(t.coef == 1 for t in self.xs + self.ys + self.q + self.r)
        desc = [t.byte() for t in self.xs + self.ys + self.q + self.r]
        limit = 10 if self.cidx is not None else 12
        assert len(desc) <= limit, "too many operand descriptors (%d)" % len(desc)
        desc += [0] * (12 - len(desc))
        w = [w0] + [int.from_bytes(bytes(desc[4 * i:4 * i + 4]), "little") for i in range(3)]
        if self.cidx is not None:
            assert 0 <= self.cidx < 65536
            w[3] |= self.cidx << 16
        return w


class Asm:
    """one routine; slots 0..NFIX-1 are the caller's registers, temporaries come from a free list"""

    def __init__(self, name, nslot, first_tmp):
        self.name, self.ins, self.nslot = name, [], nslot
        self.free_list = list(range(first_tmp, nslot))
        self.high = first_tmp

    def tmp(self, n=None):
        if n is not None:
            return [self.tmp() for _ in range(n)]
        assert self.free_list, "routine %s: out of slots" % self.name
        s = self.free_list.pop(0)
        self.high = max(self.high, s + 1)
        return s

    def free(self, *slots):
        for s in slots:
            for x in (s if isinstance(s, (list, tuple)) else [s]):
                assert x not in self.free_list
                self.free_list.insert(0, x)
        self.free_list.sort()

    # -- product class.  Posts that do not fit one instruction spill into a following LIN.
    def _prod(self, op, d, xs, ys=(), q=(), r=(), xi=False, dbl=0, **kw):
        expand = lambda ts: [T(t.slot, t.neg, t.xi, 1, t.rell, t.relv) for t in terms(ts) for _ in range(t.coef)]
        xs, ys, q, r = expand(xs), expand(ys), expand(q), expand(r)
        limit = 10 if kw.get("cidx") is not None else 12
        extra = []
        while len(xs) + len(ys) + len(q) + len(r) > limit or len(r) > 7:
            assert r and dbl == 0, "cannot split"
            extra.append(r.pop())
        self.ins.append(Ins(op, d, xs, ys, q, r, xi=xi, dbl=dbl, **kw))
        if extra:
            self.lin(d, [T(d)] + extra)

    def mul(self, d, xs, ys, q=(), r=(), xi=False, dbl=0): self._prod(OP_MUL, d, xs, ys, q, r, xi, dbl)
    def sqr(self, d, xs, q=(), r=(), xi=False, dbl=0): self._prod(OP_SQR, d, xs, (), q, r, xi, dbl)
    def mulfp(self, d, xs, sc, r=()): self._prod(OP_MULFP, d, xs, (), (), r, aux=sc)
    def mulc(self, d, xs, cidx, conjx=False, rel=False, r=()): self._prod(OP_MULC, d, xs, (), (), r, conjx=conjx, rel=rel, cidx=cidx)
    def mulfpc(self, d, cidx, sc, rel=False): self._prod(OP_MULFPC, d, (), (), (), (), rel=rel, aux=sc, cidx=cidx)

    def lin(self, d, xs, dbl=0, conj=False):
        xs = terms(xs)
        # coefficients above 4 or more than 8 terms: chain
        assert len(xs) <= 8
        self.ins.append(Ins(OP_LIN, d, xs, dbl=dbl, conjx=conj))

    def mov(self, d, s): self.lin(d, [T(s)])
    def ldc(self, d, cidx, rel=False): self.ins.append(Ins(OP_LDC, d, rel=rel, cidx=cidx))
    def ldg(self, d, field): self.ins.append(Ins(OP_LDG, d, aux=field))
    def inv(self, d, s): self.ins.append(Ins(OP_INV, d, [T(s)]))
    def fill(self, d, k, rel=False): self.ins.append(Ins(OP_FILL, d, aux=k, rel=rel))
    def spill(self, s, k, rel=False): self.ins.append(Ins(OP_SPILL, s, aux=k, rel=rel))


# ---------------------------------------------------------------------------------------- tower macros
def add_lists(a, b):
    return terms(a) + terms(b)


def fp6_mul(A, out, a, b, extra=None, out_direct_last=True):
    """out[k] = (a * b)[k] + extra[k]  (Karatsuba, 6 products; role of fp6.rs:225-242).  a[i], b[i] are term LISTS (linear
    combinations of slots), out = 3 destination slots, extra[k] = post terms.  None of out may alias a slot that a / b reads,
    except out[2], which is written by the LAST instruction."""
    extra = extra or [[], [], []]
    v0, v1, v2 = A.tmp(3)
    A.mul(v0, a[0], b[0])
    A.mul(v1, a[1], b[1])
    A.mul(v2, a[2], b[2])
    A.mul(out[0], add_lists(a[1], a[2]), add_lists(b[1], b[2]), q=[-T(v1), -T(v2)], xi=True, r=[T(v0)] + terms(extra[0]))
    A.mul(out[1], add_lists(a[0], a[1]), add_lists(b[0], b[1]), r=[-T(v0), -T(v1), T(v2).x()] + terms(extra[1]))
    A.mul(out[2], add_lists(a[0], a[2]), add_lists(b[0], b[2]), r=[-T(v0), -T(v2), T(v1)] + terms(extra[2]))
    A.free(v0, v1, v2)


def fp12_sqr(A, f):
    """f <- f^2 in place (complex squaring, fp12.rs:170-180): c1 = 2 a b, c0 = (a + b)(a + v b) - ab - v ab"""
    if getattr(A, "wide", False):
        return w_fp12_sqr(A, f)
    a, b = f[0:3], f[3:6]
    ab = A.tmp(3)
    fp6_mul(A, ab, [[T(x)] for x in a], [[T(x)] for x in b])
    w = A.tmp(3)                                            # w = ab + v ab
    A.lin(w[0], [T(ab[0]), T(ab[2]).x()])
    A.lin(w[1], [T(ab[1]), T(ab[0])])
    A.lin(w[2], [T(ab[2]), T(ab[1])])
    s = [[T(a[i]), T(b[i])] for i in range(3)]
    t = [[T(a[0]), T(b[2]).x()], [T(a[1]), T(b[0])], [T(a[2]), T(b[1])]]
    r0, r1 = A.tmp(2)
    fp6_mul(A, [r0, r1, a[2]], s, t, extra=[[-T(w[0])], [-T(w[1])], [-T(w[2])]])
    A.mov(a[0], r0)
    A.mov(a[1], r1)
    for i in range(3):
        A.lin(b[i], [T(ab[i])], dbl=1)
    A.free(ab, w, r0, r1)


def fp12_mul_by_034(A, f, l0, l3, l4):
    """f <- f * (l0 + l3 w + l4 w^3) in place: the D-twist sparse product, 13 Fp2 products (role of fp12.rs:120-129)"""
    if getattr(A, "wide", False):
        return w_fp12_mul_by_034(A, f, l0, l3, l4)
    a, b = f[0:3], f[3:6]
    # B = b * (l3 + l4 v)
    u0, u1 = A.tmp(2)
    B = A.tmp(3)
    A.mul(u0, b[0], l3)
    A.mul(u1, b[1], l4)
    A.mul(B[0], [T(b[1]), T(b[2])], l4, q=[-T(u1)], xi=True, r=[T(u0)])
    A.mul(B[1], [T(b[0]), T(b[1])], [T(l3), T(l4)], r=[-T(u0), -T(u1)])
    A.mul(B[2], [T(b[0]), T(b[2])], l3, r=[-T(u0), T(u1)])
    # s = a + b replaces b; f0' = a l0 + v B replaces a
    for i in range(3):
        A.lin(b[i], [T(a[i]), T(b[i])])
    A.mul(a[0], a[0], l0, r=[T(B[2]).x()])
    A.mul(a[1], a[1], l0, r=[T(B[0])])
    A.mul(a[2], a[2], l0, r=[T(B[1])])
    # C = s * ((l0 + l3) + l4 v);  f1' = C - f0' + v B - B
    A.mul(u0, b[0], [T(l0), T(l3)])
    A.mul(u1, b[1], l4)
    x0, x1 = A.tmp(2)
    A.mul(x0, [T(b[1]), T(b[2])], l4, q=[-T(u1)], xi=True, r=[T(u0), -T(a[0]), T(B[2]).x(), -T(B[0])])
    A.mul(x1, [T(b[0]), T(b[1])], [T(l0), T(l3), T(l4)], r=[-T(u0), -T(u1), -T(a[1]), T(B[0]), -T(B[1])])
    A.mul(b[2], [T(b[0]), T(b[2])], [T(l0), T(l3)], r=[-T(u0), T(u1), -T(a[2]), T(B[1]), -T(B[2])])
    A.mov(b[0], x0)
    A.mov(b[1], x1)
    A.free(u0, u1, B, x0, x1)


def doubling_step(A, Tp, line, sc_px, sc_py):
    """pairings.rs:901-930 on the Jacobian point Tp = (X, Y, Z) in place; line = 3 slots receiving the SCALED coefficients
    (l0 = t0 yP, l3 = t3 xP, l4 = t6: pairings.rs:888-899).  line[0] / line[1] serve as temporaries until the end."""
    if getattr(A, "wide", False):
        return w_doubling_step(A, Tp, line, sc_px, sc_py)
    X, Y, Z = Tp
    t0, t1, t3, t4, zs, q = A.tmp(6)
    t2, t5 = line[0], line[1]
    A.sqr(t0, X)
    A.sqr(t1, Y)
    A.sqr(t2, t1)
    A.sqr(t3, [T(t1), T(X)], r=[-T(t0), -T(t2)], dbl=1)
    A.lin(t4, [3 * T(t0)])
    A.sqr(t5, t4)
    A.sqr(zs, Z)
    A.lin(q, [T(t1)], dbl=2)
    A.sqr(line[2], [T(X), T(t4)], r=[-T(t0), -T(t5), -T(q)])          # t6
    A.lin(X, [T(t5), -2 * T(t3)])
    A.sqr(Z, [T(Z), T(Y)], r=[-T(t1), -T(zs)])
    A.lin(q, [T(t2)], dbl=3)
    A.mul(Y, [T(t3), -T(X)], t4, r=[-T(q)])
    A.mul(t3, [-T(t4)], zs, dbl=1)                                       # tmp3 = -2 tmp4 zsq
    A.mul(t0, Z, zs, dbl=1)                                              # tmp0 = 2 Z' zsq
    A.mulfp(line[0], t0, sc_py)
    A.mulfp(line[1], t3, sc_px)
    A.free(t0, t1, t3, t4, zs, q)


def addition_step(A, Tp, qx, qy, line, sc_px, sc_py):
    """pairings.rs:932-962; qx, qy are TERMS (so that -Q costs nothing).  The three line slots host short-lived temporaries."""
    if getattr(A, "wide", False):
        return w_addition_step(A, Tp, qx, qy, line, sc_px, sc_py)
    X, Y, Z = Tp
    qx, qy = terms(qx)[0], terms(qy)[0]
    zs, ys, a, b, c, f = A.tmp(6)
    d, e, g = line
    A.sqr(zs, Z)
    A.sqr(ys, qy)
    A.mul(a, zs, qx)                                                     # t0
    A.sqr(b, [qy, T(Z)], r=[-T(ys), -T(zs)])
    A.mul(b, b, zs)                                                      # t1
    A.lin(c, [T(a), -T(X)])                                              # t2
    A.sqr(a, c)                                                          # t3
    A.lin(d, [T(a)], dbl=2)                                              # t4
    A.mul(e, d, c)                                                       # t5
    A.lin(f, [T(b), -2 * T(Y)])                                          # t6
    A.mul(b, f, qx)                                                      # t9
    A.mul(g, d, X)                                                       # t7
    A.sqr(X, f, r=[-T(e), -2 * T(g)])
    A.sqr(Z, [T(Z), T(c)], r=[-T(zs), -T(a)])
    A.mul(a, Y, e, dbl=1)                                                # 2 Y t5
    A.mul(Y, [T(g), -T(X)], f, r=[-T(a)])
    A.sqr(a, Z)                                                          # ztsquared
    A.sqr(c, [qy, T(Z)], r=[-T(ys), -T(a)])                              # t10
    A.lin(line[2], [2 * T(b), -T(c)])                                    # 2 t9 - t10
    A.lin(a, [T(Z)], dbl=1)                                              # 2 Z'
    A.lin(c, [-T(f)], dbl=1)                                             # -2 t6
    A.mulfp(line[0], a, sc_py)
    A.mulfp(line[1], c, sc_px)
    A.free(zs, ys, a, b, c, f)


def fixed_line(A, line, cbase, sc_px, sc_py):
    """line `line counter` of the fixed point -g2 from the constant table, scaled by P0"""
    A.mulfpc(line[0], cbase + 0, sc_py, rel=True)
    A.mulfpc(line[1], cbase + 1, sc_px, rel=True)
    A.ldc(line[2], cbase + 2, rel=True)


def fp12_mul(A, x, y):
    """x <- x * y (fp12.rs:203-210, Karatsuba over Fp6, 18 products); x, y = 6 slots each"""
    if getattr(A, "wide", False):
        return w_fp12_mul(A, x, y)
    a, b, c, d = x[0:3], x[3:6], y[0:3], y[3:6]
    aa, bb = A.tmp(3), A.tmp(3)
    fp6_mul(A, aa, [[T(i)] for i in a], [[T(i)] for i in c])
    fp6_mul(A, bb, [[T(i)] for i in b], [[T(i)] for i in d])
    r0, r1 = A.tmp(2)
    fp6_mul(A, [r0, r1, b[2]], [[T(a[i]), T(b[i])] for i in range(3)], [[T(c[i]), T(d[i])] for i in range(3)],
            extra=[[-T(aa[i]), -T(bb[i])] for i in range(3)])
    A.mov(b[0], r0)
    A.mov(b[1], r1)
    A.lin(a[0], [T(aa[0]), T(bb[2]).x()])
    A.lin(a[1], [T(aa[1]), T(bb[0])])
    A.lin(a[2], [T(aa[2]), T(bb[1])])
    A.free(aa, bb, r0, r1)


def cyclotomic_square(A, f):
    """Granger-Scott squaring in the cyclotomic subgroup, in place (pairings.rs:68-115): 9 squarings + 6 linear steps"""
    if getattr(A, "wide", False):
        return w_cyclotomic_square(A, f)
    z0, z4, z3, z2, z1, z5 = f
    t0, t1, c1, t2, t3, c3 = A.tmp(6)
    # (z0, z1)
    A.sqr(t0, z0)
    A.sqr(t1, z1)
    A.sqr(c1, [T(z0), T(z1)], r=[-T(t0), -T(t1)])
    A.lin(z0, [3 * T(t1).x(), 3 * T(t0), -2 * T(z0)])
    A.lin(z1, [3 * T(c1), 2 * T(z1)])
    # (z2, z3) and (z4, z5)
    A.sqr(t0, z2)
    A.sqr(t1, z3)
    A.sqr(c1, [T(z2), T(z3)], r=[-T(t0), -T(t1)])
    A.sqr(t2, z4)
    A.sqr(t3, z5)
    A.sqr(c3, [T(z4), T(z5)], r=[-T(t2), -T(t3)])
    A.lin(z4, [3 * T(t1).x(), 3 * T(t0), -2 * T(z4)])
    A.lin(z5, [3 * T(c1), 2 * T(z5)])
    A.lin(z2, [3 * T(c3).x(), 2 * T(z2)])
    A.lin(z3, [3 * T(t3).x(), 3 * T(t2), -2 * T(z3)])
    A.free(t0, t1, c1, t2, t3, c3)


def fp6_inv(A, out, a):
    """fp6.rs:261-287; out must not alias a"""
    t0, t1, t2, d, s = A.tmp(5)
    A.mul(s, a[1], a[2])
    A.sqr(t0, a[0], r=[-T(s).x()])
    A.mul(s, a[0], a[1])
    A.sqr(t1, a[2], xi=True, r=[-T(s)])
    A.mul(s, a[0], a[2])
    A.sqr(t2, a[1], r=[-T(s)])
    A.mul(s, a[1], t2)
    A.mul(d, a[2], t1, q=[T(s)], xi=True)
    A.mul(d, a[0], t0, r=[T(d)])
    A.inv(d, d)
    A.mul(out[0], t0, d)
    A.mul(out[1], t1, d)
    A.mul(out[2], t2, d)
    A.free(t0, t1, t2, d, s)


def fp12_inv(A, f):
    """f <- 1 / f in place (fp12.rs:212-219)"""
    a, b = f[0:3], f[3:6]
    t0, t1, t2 = A.tmp(3), A.tmp(3), A.tmp(3)
    fp6_mul(A, t0, [[T(i)] for i in a], [[T(i)] for i in a])
    fp6_mul(A, t1, [[T(i)] for i in b], [[T(i)] for i in b])
    # t0 <- t0 - v t1
    A.lin(t0[0], [T(t0[0]), -T(t1[2]).x()])
    A.lin(t0[1], [T(t0[1]), -T(t1[0])])
    A.lin(t0[2], [T(t0[2]), -T(t1[1])])
    fp6_inv(A, t2, t0)
    A.free(t0)
    r = A.tmp(3)
    fp6_mul(A, r, [[T(i)] for i in a], [[T(i)] for i in t2])
    fp6_mul(A, t1, [[T(i)] for i in b], [[T(i)] for i in t2])
    for i in range(3):
        A.mov(a[i], r[i])
        A.lin(b[i], [-T(t1[i])])
    A.free(t1, t2, r)


# ---------------------------------------------------------------------------------------- constants
def build_constants():
    """VM_CONST: Fp2 constants in Montgomery form.  [0, 264): the 88 line triples of -g2; then gamma[k][i] (k = 1..3,
    i = 1..5), ENDO_U, ENDO_V, ONE."""
    from oracle import pymodel as m
    consts = []
    for c in m.line_schedule(m.g2_neg(m.G2_GEN)):
        consts.extend(c)
    idx = {"LINES": 0}
    idx["GAMMA"] = len(consts)
    for k in (1, 2, 3):
        for i in range(1, 6):
            consts.append(m.frob_gamma(k, i))
    idx["ENDO_U"] = len(consts)
    consts.append(m.ENDO_U)
    idx["ENDO_V"] = len(consts)
    consts.append(m.ENDO_V)
    idx["ONE"] = len(consts)
    consts.append((1, 0))
    return consts, idx


# ---------------------------------------------------------------------------------------- routines
F = [0, 1, 2, 3, 4, 5]           # the Fp12 accumulator: c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2
G = [6, 7, 8, 9, 10, 11]         # second Fp12 register of the final exponentiation
TP = [6, 7, 8]                   # Miller loop: the Jacobian twist point of pair 1 (X, Y, Z)
TP0 = [9, 10, 11]                # two-variable-pair loops: the twist point of pair 0
NSLOT_MILLER = 32      # assembler limit; the emitted per-family counts are the high-water marks
NSLOT_FEXP = 32


def routines(cidx, asm=None):
    R = {}
    asm = asm or Asm

    def new(name, first_tmp, nslot):
        a = asm(name, nslot, first_tmp)
        R[name] = a
        return a

    # ---- two pairs, pair 0 = (P0, -g2) through the line table, pair 1 = (P1, Q1) variable ------------------
    a = new("M2F_INIT", 9, NSLOT_MILLER)
    a.ldc(F[0], cidx["ONE"])
    for i in range(1, 6):
        a.lin(F[i], [T(F[0]), -T(F[0])])
    a.ldg(TP[0], IN_Q1X)
    a.ldg(TP[1], IN_Q1Y)
    a.ldc(TP[2], cidx["ONE"])

    a = new("M2F_DBL", 9, NSLOT_MILLER)
    fp12_sqr(a, F)
    L = a.tmp(3)
    fixed_line(a, L, cidx["LINES"], SC_P0X, SC_P0Y)
    fp12_mul_by_034(a, F, *L)
    doubling_step(a, TP, L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)

    for name, sign in (("M2F_ADDP", False), ("M2F_ADDN", True)):
        a = new(name, 9, NSLOT_MILLER)
        L = a.tmp(3)
        fixed_line(a, L, cidx["LINES"], SC_P0X, SC_P0Y)
        fp12_mul_by_034(a, F, *L)
        qx, qy = a.tmp(2)
        a.ldg(qx, IN_Q1X)
        a.ldg(qy, IN_Q1Y)
        addition_step(a, TP, T(qx), T(qy, neg=sign), L, SC_P1X, SC_P1Y)
        a.free(qx, qy)
        fp12_mul_by_034(a, F, *L)

    def psi_q(a, in_x, in_y):
        # Q1 = psi(Q): conj(x) ENDO_U, conj(y) ENDO_V (g2.rs:938-954)
        qx, qy = a.tmp(2)
        a.ldg(qx, in_x)
        a.ldg(qy, in_y)
        a.mulc(qx, qx, cidx["ENDO_U"], conjx=True)
        a.mulc(qy, qy, cidx["ENDO_V"], conjx=True)
        return qx, qy

    def psi_again(a, qx, qy):
        # Q2 = psi(Q1), in place; the caller adds -Q2
        a.mulc(qx, qx, cidx["ENDO_U"], conjx=True)
        a.mulc(qy, qy, cidx["ENDO_V"], conjx=True)

    a = new("M2F_FROB", 9, NSLOT_MILLER)
    L = a.tmp(3)
    fixed_line(a, L, cidx["LINES"], SC_P0X, SC_P0Y)            # line counter = 86
    fp12_mul_by_034(a, F, *L)
    qx, qy = psi_q(a, IN_Q1X, IN_Q1Y)
    addition_step(a, TP, T(qx), T(qy), L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)
    fixed_line(a, L, cidx["LINES"] + 3, SC_P0X, SC_P0Y)        # line 87
    fp12_mul_by_034(a, F, *L)
    psi_again(a, qx, qy)
    addition_step(a, TP, T(qx), T(qy, neg=True), L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)

    # ---- one pair (P1, Q1) ---------------------------------------------------------------------------------------
    a = new("M1_DBL", 9, NSLOT_MILLER)
    fp12_sqr(a, F)
    L = a.tmp(3)
    doubling_step(a, TP, L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)
    for name, sign in (("M1_ADDP", False), ("M1_ADDN", True)):
        a = new(name, 9, NSLOT_MILLER)
        L = a.tmp(3)
        qx, qy = a.tmp(2)
        a.ldg(qx, IN_Q1X)
        a.ldg(qy, IN_Q1Y)
        addition_step(a, TP, T(qx), T(qy, neg=sign), L, SC_P1X, SC_P1Y)
        a.free(qx, qy)
        fp12_mul_by_034(a, F, *L)
    a = new("M1_FROB", 9, NSLOT_MILLER)
    L = a.tmp(3)
    qx, qy = psi_q(a, IN_Q1X, IN_Q1Y)
    addition_step(a, TP, T(qx), T(qy), L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)
    psi_again(a, qx, qy)
    addition_step(a, TP, T(qx), T(qy, neg=True), L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)

    # ---- two variable pairs (P0, Q0), (P1, Q1): min-pk verification, twin loops ---------------------------------
    a = new("M2V_INIT", 12, NSLOT_MILLER)
    a.ldc(F[0], cidx["ONE"])
    for i in range(1, 6):
        a.lin(F[i], [T(F[0]), -T(F[0])])
    a.ldg(TP[0], IN_Q1X)
    a.ldg(TP[1], IN_Q1Y)
    a.ldc(TP[2], cidx["ONE"])
    a.ldg(TP0[0], IN_Q0X)
    a.ldg(TP0[1], IN_Q0Y)
    a.ldc(TP0[2], cidx["ONE"])
    a = new("M2V_DBL", 12, NSLOT_MILLER)
    fp12_sqr(a, F)
    L = a.tmp(3)
    doubling_step(a, TP0, L, SC_P0X, SC_P0Y)
    fp12_mul_by_034(a, F, *L)
    doubling_step(a, TP, L, SC_P1X, SC_P1Y)
    fp12_mul_by_034(a, F, *L)
    for name, sign in (("M2V_ADDP", False), ("M2V_ADDN", True)):
        a = new(name, 12, NSLOT_MILLER)
        L = a.tmp(3)
        qx, qy = a.tmp(2)
        for tp, ix, iy, sx, sy in ((TP0, IN_Q0X, IN_Q0Y, SC_P0X, SC_P0Y), (TP, IN_Q1X, IN_Q1Y, SC_P1X, SC_P1Y)):
            a.ldg(qx, ix)
            a.ldg(qy, iy)
            addition_step(a, tp, T(qx), T(qy, neg=sign), L, sx, sy)
            fp12_mul_by_034(a, F, *L)
    a = new("M2V_FROB", 12, NSLOT_MILLER)
    L = a.tmp(3)
    for tp, ix, iy, sx, sy in ((TP0, IN_Q0X, IN_Q0Y, SC_P0X, SC_P0Y), (TP, IN_Q1X, IN_Q1Y, SC_P1X, SC_P1Y)):
        # the reference order is: both pairs' first addition, then both pairs' second; products commute, so pair by pair is
        # the same field element
        qx, qy = psi_q(a, ix, iy)
        addition_step(a, tp, T(qx), T(qy), L, sx, sy)
        fp12_mul_by_034(a, F, *L)
        psi_again(a, qx, qy)
        addition_step(a, tp, T(qx), T(qy, neg=True), L, sx, sy)
        fp12_mul_by_034(a, F, *L)
        a.free(qx, qy)

    # ---- final exponentiation: routines on the Fp12 registers F (slots 0-5) and G (6-11) ------------------------
    a = new("FE_CYCSQR", 12, NSLOT_FEXP)
    cyclotomic_square(a, F)
    a = new("FE_MUL", 12, NSLOT_FEXP)            # F <- F * G
    fp12_mul(a, F, G)
    a = new("FE_SQR", 12, NSLOT_FEXP)            # F <- F^2 (general squaring)
    fp12_sqr(a, F)
    a = new("FE_CONJ", 12, NSLOT_FEXP)           # F <- conj(F)
    for i in (3, 4, 5):
        a.lin(F[i], [-T(F[i])])
    a = new("FE_INV", 6, NSLOT_FEXP)             # F <- 1 / F  (clobbers G)
    fp12_inv(a, F)
    a = new("FE_MOVGF", 12, NSLOT_FEXP)          # G <- F
    for i in range(6):
        a.mov(G[i], F[i])
    a = new("FE_MOVFG", 12, NSLOT_FEXP)          # F <- G
    for i in range(6):
        a.mov(F[i], G[i])
    # coefficient of w^i lives in: i even -> c0.c(i/2) = slot i/2, i odd -> c1.c((i-1)/2) = slot 3 + (i-1)/2
    coeff_slot = lambda i: (i // 2) if i % 2 == 0 else 3 + (i - 1) // 2
    for k in (1, 2, 3):
        a = new("FE_FROB%d" % k, 12, NSLOT_FEXP)  # F <- F^(p^k)
        for i in range(6):
            s = F[coeff_slot(i)]
            if i == 0:
                if k & 1:
                    a.lin(s, [T(s)], conj=True)
            else:
                a.mulc(s, s, cidx["GAMMA"] + 5 * (k - 1) + (i - 1), conjx=bool(k & 1))
    # scratch traffic: an Fp12 register <-> scratch variable `var` (a driver register; 6 Fp2 entries at 6 var)
    a = new("FE_LDF", 12, NSLOT_FEXP)
    for i in range(6):
        a.fill(F[i], i, rel=True)
    a = new("FE_LDG", 12, NSLOT_FEXP)
    for i in range(6):
        a.fill(G[i], i, rel=True)
    a = new("FE_STF", 12, NSLOT_FEXP)
    for i in range(6):
        a.spill(F[i], i, rel=True)
    return R


# ---------------------------------------------------------------------------------------- simulator (tests)
class Sim:
    """Executes routines on big integers (canonical Fp2 pairs).  slots: list of (c0, c1); scalars: 4 Fp; inputs: 4 Fp2;
    consts: canonical Fp2; scratch: dict"""

    def __init__(self, nslot, consts):
        self.s = [(0, 0)] * nslot
        self.consts, self.scalars, self.inputs, self.scratch, self.line, self.var = consts, [0] * 4, [(0, 0)] * 4, {}, 0, 0
        self.count = {}

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    @staticmethod
    def neg(a): return ((-a[0]) % P, (-a[1]) % P)
    @staticmethod
    def mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    @staticmethod
    def xi(a): return ((9 * a[0] - a[1]) % P, (a[0] + 9 * a[1]) % P)
    @staticmethod
    def conj(a): return (a[0], (-a[1]) % P)

    def term(self, t):
        v = self.s[t.slot]
        if t.xi:
            v = self.xi(v)
        if t.neg:
            v = self.neg(v)
        return v

    def total(self, ts):
        acc = (0, 0)
        for t in ts:
            for _ in range(t.coef):
                acc = self.add(acc, self.term(t))
        return acc

    def run(self, asm):
        for ins in asm.ins:
            self.count[ins.op] = self.count.get(ins.op, 0) + 1
            op = ins.op
            cidx = None if ins.cidx is None else ins.cidx + (3 * self.line if ins.rel else 0)
            if op in (OP_MUL, OP_SQR, OP_MULFP, OP_MULC, OP_MULFPC):
                x = self.total(ins.xs)
                if ins.conjx:
                    x = self.conj(x)
                if op == OP_MUL:
                    p = self.mul(x, self.total(ins.ys))
                elif op == OP_SQR:
                    p = self.mul(x, x)
                elif op == OP_MULFP:
                    p = (x[0] * self.scalars[ins.aux] % P, x[1] * self.scalars[ins.aux] % P)
                elif op == OP_MULC:
                    p = self.mul(x, self.consts[cidx])
                else:
                    c = self.consts[cidx]
                    p = (c[0] * self.scalars[ins.aux] % P, c[1] * self.scalars[ins.aux] % P)
                t = self.add(p, self.total(ins.q))
                if ins.xi:
                    t = self.xi(t)
                res = self.add(t, self.total(ins.r))
                for _ in range(ins.dbl):
                    res = self.add(res, res)
                self.s[ins.d] = res
            elif op == OP_LIN:
                res = self.total(ins.xs)
                for _ in range(ins.dbl):
                    res = self.add(res, res)
                if ins.conjx:
                    res = self.conj(res)
                self.s[ins.d] = res
            elif op == OP_LDC:
                self.s[ins.d] = self.consts[cidx]
            elif op == OP_LDG:
                self.s[ins.d] = self.inputs[ins.aux]
            elif op == OP_INV:
                a = self.total(ins.xs)
                n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
                self.s[ins.d] = (a[0] * n % P, (-a[1]) * n % P)
            elif op == OP_FILL:
                self.s[ins.d] = self.scratch[ins.aux + (6 * self.var if ins.rel else 0)]
            elif op == OP_SPILL:
                self.scratch[ins.aux + (6 * self.var if ins.rel else 0)] = self.s[ins.d]
            else:
                raise ValueError(op)


def naf(n):
    out = []
    while n:
        d = (2 - (n % 4)) if n & 1 else 0
        n -= d
        out.append(d)
        n >>= 1
    return out  # LSB first


X_PARAM = 0x44e992b44a6909f1
ATE_NAF = naf(6 * X_PARAM + 2)


def wnaf4(n):
    """width-4 signed digits of n, MSB first (the string k::X_WNAF4 of bn254_constants.cuh)"""
    out = []
    while n:
        if n & 1:
            d = n % 16
            if d >= 8:
                d -= 16
            n -= d
        else:
            d = 0
        out.append(d)
        n >>= 1
    return out[::-1]


def sim_miller(sim, R, fam):
    """the Miller-loop driver: fam = "M1" | "M2F" | "M2V" """
    sim.line = 0
    sim.run(R[("M2F" if fam == "M1" else fam) + "_INIT"])
    for d in reversed(ATE_NAF[:-1]):
        sim.run(R[fam + "_DBL"])
        sim.line += 1
        if d:
            sim.run(R[fam + ("_ADDP" if d > 0 else "_ADDN")])
            sim.line += 1
    sim.run(R[fam + "_FROB"])
    return [sim.s[i] for i in range(6)]


def sim_final_exp(sim, R):
    """the final-exponentiation driver on the Fp12 registers F / G and ten scratch variables (mirrors pairing.cuh:
    final_exponentiation + cyclotomic_exp_neg_x); input and result in F"""
    V_R, V_T, V_A, V_B, V_C, V_PW = 0, 1, 2, 3, 4, 5          # scratch variables; V_PW .. V_PW + 3 = the window table
    run = lambda n: sim.run(R["FE_" + n])

    def ldf(v): sim.var = v; run("LDF")
    def ldg(v): sim.var = v; run("LDG")
    def stf(v): sim.var = v; run("STF")
    digits = wnaf4(X_PARAM)

    def exp_neg_x():
        """F <- conj(F^x) with the width-4 window: table f, f^3, f^5, f^7 in scratch"""
        stf(V_PW)                      # pw[0] = f
        run("MOVGF")                   # G = f
        run("CYCSQR")                  # F = f^2
        stf(V_T)                       # t = f^2 (scratch V_T is free inside the exponentiation: see the caller)
        for i in range(1, 4):          # pw[i] = pw[i-1] * f^2
            ldf(V_PW + i - 1)
            ldg(V_T)
            run("MUL")
            stf(V_PW + i)
        ldf(V_PW + (digits[0] - 1) // 2)
        for d in digits[1:]:
            run("CYCSQR")
            if d > 0:
                ldg(V_PW + (d - 1) // 2)
                run("MUL")
            elif d < 0:
                run("CONJ")
                ldg(V_PW + (-d - 1) // 2)
                run("MUL")
                run("CONJ")
        run("CONJ")

    # easy part: t = conj(f) / f; r = frob2(t) * t
    stf(V_A)                           # a = f
    run("INV")                         # F = 1 / f   (clobbers G)
    stf(V_T)
    ldf(V_A)
    run("CONJ")
    ldg(V_T)
    run("MUL")                         # F = conj(f) / f
    stf(V_T)
    run("FROB2")
    ldg(V_T)
    run("MUL")
    stf(V_R)                           # r
    # hard part (Fuentes-Castaneda): pairing.cuh final_exponentiation
    exp_neg_x()                        # y0
    run("CYCSQR")                      # y1
    stf(V_A)                           # a = y1
    run("CYCSQR")                      # y2
    ldg(V_A)
    run("MUL")                         # y3
    stf(V_B)                           # b = y3
    exp_neg_x()                        # y4
    stf(V_C)                           # c = y4
    run("CYCSQR")                      # y5
    exp_neg_x()                        # y6
    run("CONJ")                        # f = conj(y6)
    ldg(V_C)
    run("MUL")                         # y7 = f c
    stf(V_T)
    ldf(V_B)
    run("CONJ")                        # conj(y3)
    ldg(V_T)
    run("MUL")                         # y8 = y7 conj(y3)
    stf(V_B)                           # b = y8
    ldg(V_A)
    run("MUL")                         # y9 = y8 y1
    stf(V_A)                           # a = y9  (the "f" of pairing.cuh from here on)
    ldf(V_B)
    ldg(V_C)
    run("MUL")                         # y10 = y8 y4
    ldg(V_R)
    run("MUL")                         # y11 = y10 r
    stf(V_T)
    ldf(V_A)
    run("FROB1")                       # y12
    ldg(V_T)
    run("MUL")                         # y13
    stf(V_T)
    ldf(V_B)
    run("FROB2")
    ldg(V_T)
    run("MUL")                         # y14
    stf(V_T)
    ldf(V_R)
    run("CONJ")
    ldg(V_A)
    run("MUL")                         # y15 = conj(r) y9
    run("FROB3")
    ldg(V_T)
    run("MUL")
    return [sim.s[i] for i in range(6)]


# ---------------------------------------------------------------------------------------- the WIDE machine
# Same algebra, other shape (vm.cuh: run_wide): ONE item per warp, the 16 lane pairs of the warp are 16 units that execute up
# to 16 INDEPENDENT instructions of a routine at once (a BUNDLE), exchanging values through the item's slots in shared
# memory.  It is the latency shape -- a lone pairing is a 4 ms dependent chain on one lane pair and ~0.3 ms on a warp -- used
# for small batches (one verification, the single final exponentiation of an aggregate, the per-group tail of the
# randomised path).  Differences from the sequential machine:
#   * slots are plentiful (one item per warp), so every temporary gets a FRESH slot (no reuse, no false dependencies);
#   * item scalars, item inputs, scratch variables and constants are slots too (ids below), so the only opcodes are
#     MUL, SQR, MULFP (times the c0 half of a slot), LIN and INV; 16-bit operand descriptors;
#   * a list scheduler packs each routine into bundles; inside a bundle every unit reads its operands, the warp
#     synchronises, every unit writes its result, the warp synchronises -- so an instruction may share a bundle with earlier
#     readers of its destination (write-after-read), but not with the producer of an operand or an earlier writer.
W_NOP, W_MUL, W_SQR, W_MULFP, W_LIN, W_INV = 0, 1, 2, 3, 4, 5
W_NAMES = {W_MUL: "MUL", W_SQR: "SQR", W_MULFP: "MULFP", W_LIN: "LIN", W_INV: "INV"}
W_SC, W_IN, W_VAR, W_TMP, W_CONST = 12, 16, 20, 80, 700     # slot ids: 4 scalars, 4 inputs, 60 scratch entries, temporaries, constants
W_UNITS = 16


class WIns:
    def __init__(self, op, d, xs=(), ys=(), q=(), r=(), xi=False, dbl=0, conjx=False, drelv=False):
        self.op, self.d, self.xs, self.ys, self.q, self.r = op, d, terms(xs), terms(ys), terms(q), terms(r)
        self.xi, self.dbl, self.conjx, self.drelv = xi, dbl, conjx, drelv

    def reads(self):
        return [(t.slot, t.rell, t.relv) for t in self.xs + self.ys + self.q + self.r]

    def write(self):
        return (self.d, False, self.drelv)

    def encode(self, as_mul=False):
        """as_mul: a squaring / scalar product that shares its bundle with general products is encoded as one (MUL with no Y
        terms squares X; the scalar's slot holds (s, 0), so the general product by it is the scalar product)"""
        op = W_MUL if as_mul and self.op in (W_SQR, W_MULFP) else self.op
        nx, ny, nq, nr = len(self.xs), len(self.ys), len(self.q), len(self.r)
        assert nx <= 12 and ny <= 12 and nq <= 3 and nr <= 7 and nx + ny + nq + nr <= 12 and 0 <= self.dbl <= 3 and 0 <= self.d < 1024
        if self.op != W_LIN:
            assert all(t.coef == 1 for t in self.xs + self.ys + self.q + self.r)
        assert not (self.op == W_MUL and ny == 0) and not (self.op == W_LIN and (nq or nr))
        w0 = (op | self.d << 4 | nx << 14 | ny << 18 | nq << 22 | nr << 24 | int(self.xi) << 27 | self.dbl << 28 | int(self.conjx) << 30 |
              int(self.drelv) << 31)
        desc = [t.half() for t in self.xs + self.ys + self.q + self.r]
        desc += [0] * (12 - len(desc))
        return [w0, 0] + [desc[2 * i] | desc[2 * i + 1] << 16 for i in range(6)]


class WAsm:
    """the assembler interface of the tower macros, emitting wide-machine instructions"""

    wide = True

    def __init__(self, name, nslot, first_tmp):
        self.name, self.ins, self.next_tmp, self.high = name, [], W_TMP, W_TMP

    def tmp(self, n=None):
        if n is not None:
            return [self.tmp() for _ in range(n)]
        s = self.next_tmp
        self.next_tmp += 1
        self.high = self.next_tmp
        assert self.high <= W_CONST
        return s

    def free(self, *slots):
        pass

    def _prod(self, op, d, xs, ys=(), q=(), r=(), xi=False, dbl=0, conjx=False):
        expand = lambda ts: [T(t.slot, t.neg, t.xi, 1, t.rell, t.relv) for t in terms(ts) for _ in range(t.coef)]
        xs, ys, q, r = expand(xs), expand(ys), expand(q), expand(r)
        extra = []
        while len(xs) + len(ys) + len(q) + len(r) > 12 or len(r) > 7:
            assert r and dbl == 0, "cannot split"
            extra.append(r.pop())
        self.ins.append(WIns(op, d, xs, ys, q, r, xi=xi, dbl=dbl, conjx=conjx))
        if extra:
            self.lin(d, [T(d)] + extra)

    def mul(self, d, xs, ys, q=(), r=(), xi=False, dbl=0): self._prod(W_MUL, d, xs, ys, q, r, xi, dbl)
    def sqr(self, d, xs, q=(), r=(), xi=False, dbl=0): self._prod(W_SQR, d, xs, (), q, r, xi, dbl)
    def mulfp(self, d, xs, sc, r=()): self._prod(W_MULFP, d, xs, [T(W_SC + sc)], (), r)
    def mulc(self, d, xs, cidx, conjx=False, rel=False, r=()): self._prod(W_MUL, d, xs, [T(W_CONST + cidx, rell=rel)], (), r, conjx=conjx)
    def mulfpc(self, d, cidx, sc, rel=False): self._prod(W_MULFP, d, [T(W_CONST + cidx, rell=rel)], [T(W_SC + sc)])

    def lin(self, d, xs, dbl=0, conj=False, drelv=False):
        xs = terms(xs)
        assert len(xs) <= 12
        self.ins.append(WIns(W_LIN, d, xs, dbl=dbl, conjx=conj, drelv=drelv))

    def linx(self, d, e, dbl=0):
        """d = 2^dbl (xi * sum(e.xi) + sum(e.pl)) for a linear expression e (class Lx); one multiplication by xi per instruction"""
        def split(terms_of):
            out = []
            for slot, c in terms_of.items():
                k, neg = abs(c), c < 0
                while k:
                    step = min(k, 4)
                    out.append(T(slot, neg=neg, coef=step))
                    k -= step
            return out
        xs, ys = split(e.xi), split(e.pl)
        assert xs or ys
        if not xs:
            self.ins.append(WIns(W_LIN, d, ys, dbl=dbl))
        else:
            assert len(xs) + len(ys) <= 12
            self.ins.append(WIns(W_LIN, d, xs, ys, xi=True, dbl=dbl))

    def mov(self, d, s): self.lin(d, [T(s)])
    def ldc(self, d, cidx, rel=False): self.lin(d, [T(W_CONST + cidx, rell=rel)])
    def ldg(self, d, field): self.lin(d, [T(W_IN + field)])
    def inv(self, d, s): self.ins.append(WIns(W_INV, d, [T(s)]))
    def fill(self, d, k, rel=False): self.lin(d, [T(W_VAR + k, relv=rel)])
    def spill(self, s, k, rel=False): self.lin(W_VAR + k, [T(s)], drelv=rel)


def schedule(ins, units=W_UNITS):
    """list scheduling of one routine into bundles of <= `units` instructions; returns a list of lists of indices"""
    n = len(ins)
    strict, weak = [set() for _ in ins], [set() for _ in ins]
    last_write, readers = {}, {}
    for i, I in enumerate(ins):
        for k in I.reads():
            if k in last_write:
                strict[i].add(last_write[k])              # read after write
            readers.setdefault(k, []).append(i)
        w = I.write()
        if w in last_write:
            strict[i].add(last_write[w])                  # write after write
        for j in readers.get(w, []):
            if j != i:
                weak[i].add(j)                            # write after read: same bundle is fine (reads come first)
        last_write[w] = i
        readers[w] = []
    # priority: longest chain of strict dependencies below the instruction
    succ = [[] for _ in ins]
    for i in range(n):
        for j in strict[i]:
            succ[j].append((i, 1))
        for j in weak[i]:
            succ[j].append((i, 0))
    height = [0] * n
    for i in reversed(range(n)):
        height[i] = max([height[k] + c for k, c in succ[i]], default=0)
    # A warp runs ONE code path at a time: a bundle costs the sum of the paths of the opcode classes in it.  So bundles are
    # built class by class: the class of the most urgent ready instruction first, every ready instruction of that class with
    # it; another class joins only if leaving it out would stretch the critical path (its tail is as long as what remains
    # after this bundle); linear steps always join (cheap).  Squarings and scalar products that share a bundle with general
    # products are executed as general products (the same field element, one path less).
    path = {W_MUL: 16, W_SQR: 13, W_MULFP: 12, W_LIN: 4, W_INV: 600}
    tail = [0] * n
    for i in reversed(range(n)):
        tail[i] = path[ins[i].op] + max([tail[k] for k, c in succ[i] if c], default=0)
    level, bundles, left = [None] * n, [], set(range(n))
    while left:
        b, cur = len(bundles), []

        def ready():
            return [i for i in left if all(level[j] is not None and level[j] < b for j in strict[i])
                    and all(level[j] is not None and level[j] <= b for j in weak[i])]

        def take(i):
            level[i] = b
            cur.append(i)
            left.remove(i)

        ok = ready()
        assert ok
        lead = max(ok, key=lambda i: (tail[i], -i))
        heavy = ins[lead].op != W_LIN
        classes = {ins[lead].op}
        progress = True
        while progress and len(cur) < units:
            progress = False
            for i in sorted(ready(), key=lambda i: (-tail[i], i)):
                if len(cur) >= units:
                    break
                op = ins[i].op
                product = op in (W_MUL, W_SQR, W_MULFP)
                lead_product = ins[lead].op in (W_MUL, W_SQR, W_MULFP)
                if op in classes or (product and lead_product) or (op == W_LIN and not (ins[lead].op == W_INV)):
                    take(i)
                    classes.add(op)
                    progress = True
                elif op != W_INV and ins[lead].op != W_INV and tail[i] + path[ins[lead].op] > tail[lead]:
                    take(i)
                    classes.add(op)
                    progress = True
        bundles.append(cur)
    return bundles


class Lx:
    """linear expression xi * (sum c_i s_i) + (sum c_j s_j) over slots, for WAsm.linx"""

    def __init__(self, pl=None, xi=None):
        self.pl, self.xi = dict(pl or {}), dict(xi or {})

    @staticmethod
    def of(slot, c=1):
        return Lx({slot: c})

    def __add__(self, o):
        r = Lx(self.pl, self.xi)
        for src, dst in ((o.pl, r.pl), (o.xi, r.xi)):
            for k, v in src.items():
                dst[k] = dst.get(k, 0) + v
                if dst[k] == 0:
                    del dst[k]
        return r

    def __neg__(self): return Lx({k: -v for k, v in self.pl.items()}, {k: -v for k, v in self.xi.items()})
    def __sub__(self, o): return self + (-o)
    def __rmul__(self, c): return Lx({k: c * v for k, v in self.pl.items()}, {k: c * v for k, v in self.xi.items()})

    def times_xi(self):
        assert not self.xi, "xi^2 is not representable"
        return Lx({}, self.pl)


def w_fp6(A, a, b, sparse01=False):
    """Karatsuba products of the Fp6 product a * b as FRESH slots, all independent; a[i], b[i] are term lists (b[2] = None for a
    sparse (b0, b1, 0) operand: five products).  Returns the three coefficients as Lx expressions over the product slots."""
    if sparse01:
        u0, u1, p0, p1, p2 = A.tmp(5)
        A.mul(u0, a[0], b[0])
        A.mul(u1, a[1], b[1])
        A.mul(p0, add_lists(a[1], a[2]), b[1])
        A.mul(p1, add_lists(a[0], a[1]), add_lists(b[0], b[1]))
        A.mul(p2, add_lists(a[0], a[2]), b[0])
        return [Lx({u0: 1}, {p0: 1, u1: -1}), Lx({p1: 1, u0: -1, u1: -1}), Lx({p2: 1, u0: -1, u1: 1})]
    v0, v1, v2, m0, m1, m2 = A.tmp(6)
    A.mul(v0, a[0], b[0])
    A.mul(v1, a[1], b[1])
    A.mul(v2, a[2], b[2])
    A.mul(m0, add_lists(a[1], a[2]), add_lists(b[1], b[2]))
    A.mul(m1, add_lists(a[0], a[1]), add_lists(b[0], b[1]))
    A.mul(m2, add_lists(a[0], a[2]), add_lists(b[0], b[2]))
    return [Lx({v0: 1}, {m0: 1, v1: -1, v2: -1}), Lx({m1: 1, v0: -1, v1: -1}, {v2: 1}), Lx({m2: 1, v0: -1, v2: -1, v1: 1})]


def times_v(c):
    """v * (c0, c1, c2) = (xi c2, c0, c1)"""
    return [c[2].times_xi(), c[0], c[1]]


def w_fp12_sqr(A, f):
    """f <- f^2 (complex squaring): twelve independent products, then six linear steps straight into f"""
    a, b = f[0:3], f[3:6]
    ab = w_fp6(A, [[T(x)] for x in a], [[T(x)] for x in b])
    s = [[T(a[i]), T(b[i])] for i in range(3)]
    t = [[T(a[0]), T(b[2]).x()], [T(a[1]), T(b[0])], [T(a[2]), T(b[1])]]
    st = w_fp6(A, s, t)
    vab = times_v(ab)
    for i in range(3):
        A.linx(a[i], st[i] - ab[i] - vab[i])
        A.linx(b[i], ab[i], dbl=1)


def w_fp12_mul_by_034(A, f, l0, l3, l4):
    """f <- f * (l0 + l3 w + l4 w^3): thirteen independent products, six linear steps"""
    a, b = f[0:3], f[3:6]
    A0 = A.tmp(3)
    for i in range(3):
        A.mul(A0[i], a[i], l0)
    B = w_fp6(A, [[T(x)] for x in b], [[T(l3)], [T(l4)], None], sparse01=True)
    C = w_fp6(A, [[T(a[i]), T(b[i])] for i in range(3)], [[T(l0), T(l3)], [T(l4)], None], sparse01=True)
    vB = times_v(B)
    for i in range(3):
        A.linx(b[i], C[i] - Lx.of(A0[i]) - B[i])
    for i in range(3):
        A.linx(a[i], Lx.of(A0[i]) + vB[i])


def w_fp12_mul(A, x, y):
    """x <- x * y: eighteen independent products, six linear steps"""
    a, b, c, d = x[0:3], x[3:6], y[0:3], y[3:6]
    aa = w_fp6(A, [[T(i)] for i in a], [[T(i)] for i in c])
    bb = w_fp6(A, [[T(i)] for i in b], [[T(i)] for i in d])
    ss = w_fp6(A, [[T(a[i]), T(b[i])] for i in range(3)], [[T(c[i]), T(d[i])] for i in range(3)])
    vbb = times_v(bb)
    for i in range(3):
        A.linx(b[i], ss[i] - aa[i] - bb[i])
    for i in range(3):
        A.linx(a[i], aa[i] + vbb[i])


def w_cyclotomic_square(A, f):
    """Granger-Scott squaring: nine independent squarings, six linear steps (pairings.rs:68-115)"""
    z0, z4, z3, z2, z1, z5 = f
    t0, t1, c1, u0, u1, d1, t2, t3, c3 = A.tmp(9)
    A.sqr(t0, z0)
    A.sqr(t1, z1)
    A.sqr(c1, [T(z0), T(z1)])
    A.sqr(u0, z2)
    A.sqr(u1, z3)
    A.sqr(d1, [T(z2), T(z3)])
    A.sqr(t2, z4)
    A.sqr(t3, z5)
    A.sqr(c3, [T(z4), T(z5)])
    A.linx(z0, Lx({t0: 3, z0: -2}, {t1: 3}))
    A.linx(z1, Lx({c1: 3, t0: -3, t1: -3, z1: 2}))
    A.linx(z4, Lx({u0: 3, z4: -2}, {u1: 3}))
    A.linx(z5, Lx({d1: 3, u0: -3, u1: -3, z5: 2}))
    A.linx(z2, Lx({z2: 2}, {c3: 3, t2: -3, t3: -3}))
    A.linx(z3, Lx({t2: 3, z3: -2}, {t3: 3}))


def w_doubling_step(A, Tp, line, sc_px, sc_py):
    """pairings.rs:901-930 with the dependent chain cut to three product levels: every product is written over values that are
    already there (X, Y, Z and products of earlier levels), the glue is linear steps"""
    X, Y, Z = Tp
    t0, t1, zs, zy, t2, t3p, t5, t6p, w3, w0, u, yp = A.tmp(12)
    A.sqr(t0, X)
    A.sqr(t1, Y)
    A.sqr(zs, Z)
    A.sqr(zy, [T(Z), T(Y)])
    A.sqr(t2, t1)
    A.sqr(t3p, [T(t1), T(X)])
    A.sqr(t5, [3 * T(t0)])
    A.sqr(t6p, [T(X), 3 * T(t0)])
    A.mul(w3, t0, zs)                                                   # tmp3 = -6 t0 zs
    A.mul(w0, [T(zy), -T(t1), -T(zs)], zs)                              # Z' zs
    A.linx(u, Lx({t3p: 1, t0: -1, t2: -1}))                             # t3 = 2 u
    A.mul(yp, [6 * T(u), -T(t5)], [3 * T(t0)])                          # (t3 - X') t4
    A.mulfp(line[0], [2 * T(w0)], sc_py)
    A.mulfp(line[1], [-6 * T(w3)], sc_px)
    A.linx(line[2], Lx({t6p: 1, t0: -1, t5: -1, t1: -4}))
    A.linx(X, Lx({t5: 1, u: -4}))
    A.linx(Z, Lx({zy: 1, t1: -1, zs: -1}))
    A.linx(Y, Lx({yp: 1, t2: -8}))


def w_addition_step(A, Tp, qx, qy, line, sc_px, sc_py):
    """pairings.rs:932-962, five product levels; qx, qy are TERMS"""
    X, Y, Z = Tp
    qx, qy = terms(qx)[0], terms(qy)[0]
    zs, ys, bp, a, b, a3, b9, fp, zcp, e, g, zt, t10p, ye, yp = A.tmp(15)
    A.sqr(zs, Z)
    A.sqr(ys, qy)
    A.sqr(bp, [qy, T(Z)])
    A.mul(a, zs, qx)                                                    # t0
    A.mul(b, [T(bp), -T(ys), -T(zs)], zs)                               # t1
    c = [T(a), -T(X)]                                                   # t2
    f = [T(b), -2 * T(Y)]                                               # t6
    A.sqr(a3, c)                                                        # t3
    A.mul(b9, f, qx)                                                    # t9
    A.sqr(fp, f)
    A.sqr(zcp, [T(Z)] + c)                                              # (Z + t2)^2
    A.mul(e, [4 * T(a3)], c)                                            # t5
    A.mul(g, [4 * T(a3)], X)                                            # t7
    zn = [T(zcp), -T(zs), -T(a3)]                                       # Z'
    A.sqr(zt, zn)
    A.sqr(t10p, [qy] + zn)
    A.mulfp(line[0], [2 * t for t in zn], sc_py)
    A.mulfp(line[1], [-2 * T(b), 4 * T(Y)], sc_px)
    A.mul(ye, Y, e)
    A.mul(yp, [3 * T(g), T(e), -T(fp)], f)                              # (t7 - X') t6
    A.linx(line[2], Lx({b9: 2, t10p: -1, ys: 1, zt: 1}))
    A.linx(X, Lx({fp: 1, e: -1, g: -2}))
    A.linx(Z, Lx({zcp: 1, zs: -1, a3: -1}))
    A.linx(Y, Lx({yp: 1, ye: -2}))


class WSim(Sim):
    """bundle-level simulator of the wide machine: all units read, then all units write"""

    def __init__(self, consts):
        Sim.__init__(self, W_CONST, consts)
        self.s = {}

    def set_item(self, scalars, inputs):
        for i, v in enumerate(scalars):
            self.s[W_SC + i] = (v, 0)
        for i, v in enumerate(inputs):
            self.s[W_IN + i] = v

    def addr(self, slot, rell, relv):
        return slot + (3 * self.line if rell else 0) + (6 * self.var if relv else 0)

    def term(self, t):
        a = self.addr(t.slot, t.rell, t.relv)
        v = self.consts[a - W_CONST] if a >= W_CONST else self.s[a]
        if t.xi:
            v = self.xi(v)
        if t.neg:
            v = self.neg(v)
        return v

    def result(self, ins):
        if ins.op == W_LIN:
            res = self.total(ins.xs)
            if ins.xi:
                res = self.xi(res)
            res = self.add(res, self.total(ins.ys))
            for _ in range(ins.dbl):
                res = self.add(res, res)
            return self.conj(res) if ins.conjx else res
        x = self.total(ins.xs)
        if ins.op == W_INV:
            n = pow((x[0] * x[0] + x[1] * x[1]) % P, P - 2, P)
            return (x[0] * n % P, (-x[1]) * n % P)
        if ins.conjx:
            x = self.conj(x)
        if ins.op == W_MUL:
            p = self.mul(x, self.total(ins.ys))
        elif ins.op == W_SQR:
            p = self.mul(x, x)
        else:
            sc = self.total(ins.ys)[0]
            p = (x[0] * sc % P, x[1] * sc % P)
        t = self.add(p, self.total(ins.q))
        if ins.xi:
            t = self.xi(t)
        res = self.add(t, self.total(ins.r))
        for _ in range(ins.dbl):
            res = self.add(res, res)
        return res

    def run(self, asm):
        for b in asm.bundles:
            out = [(self.addr(*asm.ins[i].write()), self.result(asm.ins[i])) for i in b]
            assert len({a for a, _ in out}) == len(out)
            for a, v in out:
                self.s[a] = v
            self.count[len(b)] = self.count.get(len(b), 0) + 1


def split_posts(a):
    """A product whose post terms (q, r) become available later than its factors is split into the bare product (fresh slot)
    and a linear step, so that the products of a Karatsuba wave do not wait for one another."""
    avail, out = {}, []
    ready = lambda ts: max([avail.get((t.slot, t.rell, t.relv), 0) for t in ts], default=0)
    for I in a.ins:
        if I.op in (W_MUL, W_SQR, W_MULFP) and (I.q or I.r) and ready(I.q + I.r) > ready(I.xs + I.ys):
            assert not any(t.xi for t in I.q) or not I.xi
            t = a.tmp()
            bare = WIns(I.op, t, I.xs, I.ys, conjx=I.conjx)
            flip = lambda u: T(u.slot, u.neg, u.xi or I.xi, u.coef, u.rell, u.relv)
            lin = WIns(W_LIN, I.d, [T(t, xi=I.xi)] + [flip(u) for u in I.q] + list(I.r), dbl=I.dbl)
            lvl = ready(bare.xs + bare.ys)
            avail[bare.write()] = lvl + 1
            avail[lin.write()] = max(lvl + 1, ready(lin.xs)) + 1
            out += [bare, lin]
        else:
            avail[I.write()] = ready(I.xs + I.ys + I.q + I.r) + 1
            out.append(I)
    a.ins = out


def sink_into_moves(a):
    """`t = expr; ...; d = t` (the sequential macros protect in-place updates that way) becomes `d = expr` at the position of
    the move when nothing in between touches t or changes an operand of expr: one bundle less on the critical path."""
    key = lambda t: (t.slot, t.rell, t.relv)
    changed = True
    while changed:
        changed = False
        for i, I in enumerate(a.ins):
            if not (I.op == W_LIN and len(I.xs) == 1 and not I.dbl and not I.conjx and not I.drelv):
                continue
            t = I.xs[0]
            if t.neg or t.xi or t.coef != 1 or t.rell or t.relv or t.slot < W_TMP:
                continue
            src = key(t)
            js = [j for j in range(i) if a.ins[j].write() == src]
            if not js:
                continue
            j = js[-1]
            J = a.ins[j]
            users = [k for k in range(j + 1, len(a.ins)) if k != i and src in a.ins[k].reads()]
            later_writers = [k for k in range(j + 1, len(a.ins)) if a.ins[k].write() == src]
            if users and not (later_writers and later_writers[0] < users[0]):
                continue
            operands = set(J.reads())
            if any(a.ins[k].write() in operands for k in range(j + 1, i)):
                continue
            a.ins[i] = WIns(J.op, I.d, J.xs, J.ys, J.q, J.r, xi=J.xi, dbl=J.dbl, conjx=J.conjx)
            del a.ins[j]
            changed = True
            break


def routines_wide(cidx):
    R = routines(cidx, WAsm)
    for a in R.values():
        sink_into_moves(a)
        split_posts(a)
        a.bundles = schedule(a.ins)
    return R


def emit_wide():
    consts, cidx = build_constants()
    R = routines_wide(cidx)
    out = ["// GENERATED by tools/gen_vm_programs.py -- do not edit.  Bundled programs of the WIDE slot machine (vm.cuh: run_wide): one item",
           "// per warp, up to 16 independent instructions per bundle.  Routine ids and constants are those of vm_prog.cuh.",
           "#pragma once", "#include <stdint.h>", "", "namespace bn {", "namespace vm {", ""]
    for name, val in (("W_MUL", W_MUL), ("W_SQR", W_SQR), ("W_MULFP", W_MULFP), ("W_LIN", W_LIN), ("W_INV", W_INV), ("W_SC", W_SC), ("W_IN", W_IN),
                      ("W_VAR", W_VAR), ("W_TMP", W_TMP), ("W_CONST", W_CONST), ("W_UNITS", W_UNITS)):
        out.append("constexpr int %s = %d;" % (name, val))
    out.append("constexpr int W_NSLOT = %d;   // slots per item (high-water mark of the temporaries)" % max(a.high for a in R.values()))
    words, boff, dirs = [], [0], []
    for n in R:
        a = R[n]
        dirs.append((len(boff) - 1, len(a.bundles)))
        for b in a.bundles:
            mixed = len({a.ins[i].op for i in b if a.ins[i].op in (W_MUL, W_SQR, W_MULFP)}) > 1
            for i in b:
                words.append(a.ins[i].encode(as_mul=mixed))
            boff.append(len(words))
    out.append("// routine directory: {first bundle, bundle count}; bundle b = instructions [W_BOFF[b], W_BOFF[b + 1])")
    out.append("BN_GDECL uint16_t W_DIR[R_COUNT][2] = {" + ", ".join("{%d, %d}" % d for d in dirs) + "};")
    out.append("BN_GDECL uint16_t W_BOFF[%d] = {%s};" % (len(boff), ", ".join(str(x) for x in boff)))
    out.append("// %d instructions in %d bundles, 32 bytes each" % (len(words), len(boff) - 1))
    out.append("BN_GDECL uint32_t W_PROG[%d][8] = {" % len(words))
    for w in words:
        out.append("    {" + ", ".join("0x%08xu" % x for x in w) + "},")
    out.append("};")
    out += ["", "}  // namespace vm", "}  // namespace bn", ""]
    (ROOT / "bls_bn254_b200" / "csrc" / "vmw_prog.cuh").write_text("\n".join(out))
    return R


# ---------------------------------------------------------------------------------------- emission
def mont_words(v):
    v = v * R256 % P
    return ", ".join("0x%08xu" % ((v >> (32 * i)) & 0xffffffff) for i in range(8))


def emit():
    consts, cidx = build_constants()
    R = routines(cidx)
    order = list(R)
    out = ["// GENERATED by tools/gen_vm_programs.py -- do not edit.  Programs, constants and routine directory of the pairing slot",
           "// machine (vm.cuh).  Instruction format and semantics: tools/gen_vm_programs.py (ISA section) and vm.cuh.",
           "#pragma once", "#include <stdint.h>", "", "namespace bn {", "namespace vm {", ""]
    for name, val in (("OP_MUL", OP_MUL), ("OP_SQR", OP_SQR), ("OP_MULFP", OP_MULFP), ("OP_MULC", OP_MULC), ("OP_MULFPC", OP_MULFPC),
                      ("OP_LIN", OP_LIN), ("OP_LDC", OP_LDC), ("OP_LDG", OP_LDG), ("OP_INV", OP_INV), ("OP_FILL", OP_FILL), ("OP_SPILL", OP_SPILL)):
        out.append("constexpr int %s = %d;" % (name, val))
    hw = lambda pre: max(a.high for n, a in R.items() if n.startswith(pre))
    out.append("// slots per lane each kernel family needs (high-water marks of its routines)")
    out.append("constexpr int NSLOT_M1 = %d;    // one-pair loops and the fixed-table two-pair loop" % max(hw("M1_"), hw("M2F_")))
    out.append("constexpr int NSLOT_M2V = %d;   // two variable pairs" % hw("M2V_"))
    out.append("constexpr int NSLOT_FEXP = %d;  // final exponentiation" % hw("FE_"))
    out.append("constexpr int NCONST = %d;" % len(consts))
    out.append("constexpr int FEXP_SCRATCH_FP2 = 60;   // per-item scratch of the final exponentiation: 10 Fp12 variables")
    out.append("")
    out.append("enum Routine : int {")
    for n in order:
        out.append("    R_%s," % n)
    out.append("    R_COUNT")
    out.append("};")
    words, dirs, pos = [], [], 0
    for n in order:
        enc = [w for ins in R[n].ins for w in ins.encode()]
        dirs.append((pos, len(R[n].ins)))
        words += enc
        pos += len(R[n].ins)
    out.append("// routine directory: {first instruction, instruction count}")
    out.append("BN_KDECL uint16_t VM_DIR[R_COUNT][2] = {" + ", ".join("{%d, %d}" % d for d in dirs) + "};")
    out.append("// %d instructions, 16 bytes each" % pos)
    out.append("BN_KDECL uint32_t VM_PROG[%d][4] = {" % pos)
    for i in range(pos):
        out.append("    {0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}," % tuple(words[4 * i:4 * i + 4]))
    out.append("};")
    out.append("// Fp2 constants, Montgomery form: [0, 264) the 88 line triples of -g2, then gamma[k][i], ENDO_U, ENDO_V, ONE")
    out.append("BN_KDECL uint32_t VM_CONST[NCONST][2][8] = {")
    for c in consts:
        out.append("    {{%s}, {%s}}," % (mont_words(c[0]), mont_words(c[1])))
    out.append("};")
    out += ["", "}  // namespace vm", "}  // namespace bn", ""]
    path = ROOT / "bls_bn254_b200" / "csrc" / "vm_prog.cuh"
    path.write_text("\n".join(out))
    return R, consts, cidx, pos


if __name__ == "__main__":
    RW = emit_wide()
    print("vmw_prog.cuh: %d routines, %d instructions in %d bundles, %d slots per item" % (
        len(RW), sum(len(a.ins) for a in RW.values()), sum(len(a.bundles) for a in RW.values()), max(a.high for a in RW.values())))
    for name, a in RW.items():
        heavy = lambda b: any(a.ins[i].op != W_LIN for i in b)
        print("  %-10s %3d instr in %3d bundles (%2d with products)  %s" % (name, len(a.ins), len(a.bundles), sum(heavy(b) for b in a.bundles),
                                                                          " ".join(("%d" if heavy(b) else "(%d)") % len(b) for b in a.bundles)))
    R, consts, cidx, n = emit()
    print("vm_prog.cuh: %d routines, %d instructions (%d bytes), %d constants" % (len(R), n, 16 * n, len(consts)))
    for name, a in R.items():
        if True:
            kinds = {}
            for i in a.ins:
                kinds[OPNAMES[i.op]] = kinds.get(OPNAMES[i.op], 0) + 1
            print("  %-10s %3d instr, slots < %2d  %s" % (name, len(a.ins), a.high, kinds))
