/* ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement of the reference's Goldilocks field and its quadratic extension.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.  The product path (olavm_b200/) never links or calls it.
 *
 * Follows:
 *   plonky2/field/src/goldilocks_field.rs      (ORDER :126, EPSILON :14, add :191-213, sub :228-250,
 *                                               mul :259-266, reduce128 :342-355, to_canonical :162-170,
 *                                               generators :66-77)
 *   plonky2/field/src/types.rs                 (primitive_root_of_unity :240-244, coset_shift :430,
 *                                               exp_u64, exp_power_of_2)
 *   plonky2/field/src/goldilocks_extensions.rs (W = 7 :19, quadratic mul :120-…)
 *   plonky2/field/src/extension/quadratic.rs   (add/sub/mul/scalar_mul/inverse)
 *
 * Deviation by design: the reference keeps non-canonical u64 representatives in memory and
 * canonicalises on compare/serialise; every function here returns the canonical representative
 * in [0, p).  The field element denoted is identical, which is all that reaches a proof byte.
 */
#ifndef ORC_GL_H
#define ORC_GL_H
#include <stdint.h>
#include <stddef.h>

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL
/* goldilocks_field.rs:66 MULTIPLICATIVE_GROUP_GENERATOR, :77 POWER_OF_TWO_GENERATOR (order 2^32) */
#define GL_GEN 7ULL
#define GL_TWO_ADIC_GEN 1753635133440165772ULL
#define GL_TWO_ADICITY 32

typedef unsigned __int128 u128;

/* (the conditional corrections are written as masks: on random field elements the branches are coin flips, and the
 * mispredictions, not the arithmetic, were most of the oracle's run time) */
static inline uint64_t gl_canon(uint64_t x) { return x - ((0 - (uint64_t)(x >= GL_P)) & GL_P); }

static inline uint64_t gl_add(uint64_t a, uint64_t b) {
    /* a, b canonical */
    uint64_t s = a + b;
    return s - ((0 - (uint64_t)((s < a) | (s >= GL_P))) & GL_P);
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return (a - b) + ((0 - (uint64_t)(a < b)) & GL_P); }
static inline uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

/* goldilocks_field.rs:342-355 reduce128, then canonicalised */
static inline uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    t0 -= (0 - (uint64_t)(lo < hi_hi)) & GL_EPS;
    uint64_t t1 = hi_lo * GL_EPS;
    uint64_t t2 = t0 + t1;
    t2 += (0 - (uint64_t)(t2 < t0)) & GL_EPS;
    return gl_canon(t2);
}
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
static inline uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }

static inline uint64_t gl_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}
/* unique inverse; the reference uses a binary-GCD variant (inversion.rs), the value is the same */
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }

/* types.rs:240-244 */
static inline uint64_t gl_root_of_unity(int n_log) {
    uint64_t b = GL_TWO_ADIC_GEN;
    for (int i = 0; i < GL_TWO_ADICITY - n_log; i++) b = gl_sqr(b);
    return b;
}

/* ---- quadratic extension F[X]/(X^2 - 7), element = (c0, c1) ---- */
typedef struct {
    uint64_t c0, c1;
} gl2_t;
#define GL2_W 7ULL

static inline gl2_t gl2_make(uint64_t a, uint64_t b) {
    gl2_t r = {a, b};
    return r;
}
static inline gl2_t gl2_from_base(uint64_t a) { return gl2_make(a, 0); }
static inline gl2_t gl2_add(gl2_t a, gl2_t b) { return gl2_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
static inline gl2_t gl2_sub(gl2_t a, gl2_t b) { return gl2_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline gl2_t gl2_neg(gl2_t a) { return gl2_make(gl_neg(a.c0), gl_neg(a.c1)); }
/* goldilocks_extensions.rs ext2_mul: (a0 b0 + W a1 b1, a0 b1 + a1 b0) */
static inline gl2_t gl2_mul(gl2_t a, gl2_t b) {
    uint64_t c0 = gl_add(gl_mul(a.c0, b.c0), gl_mul(GL2_W, gl_mul(a.c1, b.c1)));
    uint64_t c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return gl2_make(c0, c1);
}
static inline gl2_t gl2_scalar_mul(gl2_t a, uint64_t s) { return gl2_make(gl_mul(a.c0, s), gl_mul(a.c1, s)); }
static inline int gl2_eq(gl2_t a, gl2_t b) { return a.c0 == b.c0 && a.c1 == b.c1; }
/* extension/quadratic.rs try_inverse: a^-1 = conj(a) / (a0^2 - W a1^2) */
static inline gl2_t gl2_inv(gl2_t a) {
    uint64_t n = gl_sub(gl_sqr(a.c0), gl_mul(GL2_W, gl_sqr(a.c1)));
    uint64_t ni = gl_inv(n);
    return gl2_make(gl_mul(a.c0, ni), gl_mul(gl_neg(a.c1), ni));
}
static inline gl2_t gl2_pow(gl2_t b, uint64_t e) {
    gl2_t r = gl2_make(1, 0);
    while (e) {
        if (e & 1) r = gl2_mul(r, b);
        b = gl2_mul(b, b);
        e >>= 1;
    }
    return r;
}

static inline uint32_t orc_log2_strict(size_t n) {
    uint32_t l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}
/* cfft/mod.rs:282-290 permute_index == bit reversal on log2(size) bits */
static inline size_t orc_bitrev(size_t x, uint32_t bits) {
    size_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
#endif
