/* TEST INFRASTRUCTURE (oracle) ONLY -- included twice by bn254_oracle.c, once per group.
 *
 * Homogeneous-projective complete formulas of Renes-Costello-Batina (eprint 2015/1060,
 * Alg. 7 and 9 for a = 0), which is what the reference uses:
 *   G1: /root/reference/src/inner_types/g1.rs:744-786 (addition), 789-818 (double)
 *   G2: /root/reference/src/inner_types/g2.rs:789-831 (addition), 834-863 (double)
 * plus projective -> affine (g1.rs:104-118, g2.rs:98-112) and MSB-first double-and-add
 * (g1.rs:821-841, g2.rs:866-886; same group element, variable time here).
 *
 * Parameters: FE (field type), FN(x) (field function prefix), PT (point type), PN(x) (point
 * function prefix), MUL_BY_3B(r,a).
 */

static void PN(set_identity)(PT *r) {
    FN(set_zero)(&r->x);
    FN(set_one)(&r->y);
    FN(set_zero)(&r->z);
}

static int PN(is_identity)(const PT *a) { return FN(is_zero)(&a->z); }

static void PN(add)(PT *r, const PT *a, const PT *b) {
    FE t0, t1, t2, t3, t4, x3, y3, z3;
    FN(mul)(&t0, &a->x, &b->x);
    FN(mul)(&t1, &a->y, &b->y);
    FN(mul)(&t2, &a->z, &b->z);
    FN(add)(&t3, &a->x, &a->y);
    FN(add)(&t4, &b->x, &b->y);
    FN(mul)(&t3, &t3, &t4);
    FN(add)(&t4, &t0, &t1);
    FN(sub)(&t3, &t3, &t4);
    FN(add)(&t4, &a->y, &a->z);
    FN(add)(&x3, &b->y, &b->z);
    FN(mul)(&t4, &t4, &x3);
    FN(add)(&x3, &t1, &t2);
    FN(sub)(&t4, &t4, &x3);
    FN(add)(&x3, &a->x, &a->z);
    FN(add)(&y3, &b->x, &b->z);
    FN(mul)(&x3, &x3, &y3);
    FN(add)(&y3, &t0, &t2);
    FN(sub)(&y3, &x3, &y3);
    FN(add)(&x3, &t0, &t0);
    FN(add)(&t0, &x3, &t0);
    MUL_BY_3B(&t2, &t2);
    FN(add)(&z3, &t1, &t2);
    FN(sub)(&t1, &t1, &t2);
    MUL_BY_3B(&y3, &y3);
    FN(mul)(&x3, &t4, &y3);
    FN(mul)(&t2, &t3, &t1);
    FN(sub)(&x3, &t2, &x3);
    FN(mul)(&y3, &y3, &t0);
    FN(mul)(&t1, &t1, &z3);
    FN(add)(&y3, &t1, &y3);
    FN(mul)(&t0, &t0, &t3);
    FN(mul)(&z3, &z3, &t4);
    FN(add)(&z3, &z3, &t0);
    r->x = x3; r->y = y3; r->z = z3;
}

static void PN(dbl)(PT *r, const PT *a) {
    FE t0, t1, t2, x3, y3, z3;
    if (PN(is_identity)(a)) { PN(set_identity)(r); return; }
    FN(sqr)(&t0, &a->y);
    FN(add)(&z3, &t0, &t0);
    FN(add)(&z3, &z3, &z3);
    FN(add)(&z3, &z3, &z3);
    FN(mul)(&t1, &a->y, &a->z);
    FN(sqr)(&t2, &a->z);
    MUL_BY_3B(&t2, &t2);
    FN(mul)(&x3, &t2, &z3);
    FN(add)(&y3, &t0, &t2);
    FN(mul)(&z3, &t1, &z3);
    FN(add)(&t1, &t2, &t2);
    FN(add)(&t2, &t1, &t2);
    FN(sub)(&t0, &t0, &t2);
    FN(mul)(&y3, &t0, &y3);
    FN(add)(&y3, &x3, &y3);
    FN(mul)(&t1, &a->x, &a->y);
    FN(mul)(&x3, &t0, &t1);
    FN(add)(&x3, &x3, &x3);
    r->x = x3; r->y = y3; r->z = z3;
}

/* affine (x, y, inf) <- projective.  Identity -> (0, 1), inf = 1 (g1.rs:265-271). */
static void PN(to_affine)(FE *x, FE *y, int *inf, const PT *a) {
    if (PN(is_identity)(a)) {
        FN(set_zero)(x); FN(set_one)(y); *inf = 1;
        return;
    }
    FE zi;
    FN(inv)(&zi, &a->z);
    FN(mul)(x, &a->x, &zi);
    FN(mul)(y, &a->y, &zi);
    *inf = 0;
}

static void PN(from_affine)(PT *r, const FE *x, const FE *y, int inf) {
    if (inf) { PN(set_identity)(r); return; }
    r->x = *x; r->y = *y; FN(set_one)(&r->z);
}

/* r = [k]a, k given as nbits-bit big integer in little-endian 64-bit limbs; MSB first. */
static void PN(mul_limbs)(PT *r, const PT *a, const uint64_t *k, int nbits) {
    PT acc, base = *a;
    PN(set_identity)(&acc);
    for (int i = nbits - 1; i >= 0; i--) {
        PN(dbl)(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) PN(add)(&acc, &acc, &base);
    }
    *r = acc;
}

/* r = [a + b lambda]p by one joint MSB-first double-and-add over {p, phi(p), p + phi(p)}, phi(x, y) = (beta x, y);
 * a, b < 2^nbits (nbits <= 64).  Randomised batch verification only (no counterpart in the reference). */
static void PN(mul_glv)(PT *r, const PT *p, uint64_t a, uint64_t b, int nbits) {
    PT tab[3], acc;
    tab[0] = *p;
    tab[1] = *p;
    FN(glv_beta)(&tab[1].x, &p->x);
    PN(add)(&tab[2], &tab[0], &tab[1]);
    PN(set_identity)(&acc);
    for (int i = nbits - 1; i >= 0; i--) {
        PN(dbl)(&acc, &acc);
        unsigned sel = (unsigned)((a >> i) & 1) | (unsigned)(((b >> i) & 1) << 1);
        if (sel) PN(add)(&acc, &acc, &tab[sel - 1]);
    }
    *r = acc;
}
