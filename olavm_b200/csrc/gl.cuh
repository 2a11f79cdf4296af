// Goldilocks field arithmetic for sm_100a (p = 2^64 - 2^32 + 1) and its quadratic extension.
//
// Replaces the reference's CPU field on the proving hot path:
//   plonky2/field/src/goldilocks_field.rs  (add :191-213, sub :228-250, mul :259-266, reduce128 :342-355)
//   plonky2/field/src/goldilocks_extensions.rs :14-39 (X^2 - 7), extension/quadratic.rs
//
// Representation: every value stored in HBM or returned by these functions is the CANONICAL
// representative in [0, p).  (The reference keeps lazy representatives in [0, 2^64) and only
// canonicalises on compare/serialise, goldilocks_field.rs:162-170; the field element is the same.)
// All arithmetic is 64-bit integer math on the INT32/IMAD pipes -- there is no tensor-core path for
// modular arithmetic.  The same header compiles for the host (challenger, twiddle setup).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#define GL_D __device__ __forceinline__
#else
#define GL_HD inline
#define GL_D inline
#endif

namespace gl {

static constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
static constexpr uint64_t EPS = 0xFFFFFFFFULL;
static constexpr uint64_t GEN = 7;                              // goldilocks_field.rs:66
static constexpr uint64_t TWO_ADIC_GEN = 1753635133440165772ULL;  // goldilocks_field.rs:77 (order 2^32)
static constexpr int TWO_ADICITY = 32;

GL_HD uint64_t canon(uint64_t x) { return x >= P ? x - P : x; }

// a, b canonical -> canonical
GL_HD uint64_t add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    uint64_t t = s + EPS;  // s - p (mod 2^64); wraps iff s >= p
    return ((s < a) | (t < s)) ? t : s;
}
GL_HD uint64_t sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return (a < b) ? d - EPS : d;  // d + p (mod 2^64)
}
GL_HD uint64_t neg(uint64_t a) { return a ? P - a : 0; }
GL_HD uint64_t dbl(uint64_t a) { return add(a, a); }

// x = hi*2^64 + lo  ->  canonical x mod p.  2^64 = 2^32 - 1, 2^96 = -1 (mod p).
GL_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= EPS;
    uint64_t t1 = (hi_lo << 32) - hi_lo;  // hi_lo * (2^32 - 1)
    uint64_t t2 = t0 + t1;
    if (t2 < t1) t2 += EPS;
    return canon(t2);
}

GL_HD void mul_wide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

// ---- device fast path ------------------------------------------------------------------------------
// "lazy" values are arbitrary u64 representatives in [0, 2^64); "canonical" ones are < p.
//   mul_lazy(any, any)          -> lazy        16 SASS instructions (7 for the 128-bit product, 9 to reduce)
//   canon_fast(lazy)            -> canonical    4
//   add_lc(lazy, canonical)     -> lazy         5      (a + v < 2^64 + p  =>  a single +EPS correction suffices)
//   sub_lc(lazy, canonical)     -> lazy         5
//   mad_lazy(any, any, any)     -> lazy        a*b + c reduced once (Field::multiply_accumulate,
//                                              goldilocks_field.rs:110-113)
// How the counts are reached (checked with cuobjdump -sass, tools/microbench/prims.cu): the 128-bit product is left to
// the compiler (4 IMAD.WIDE chained through their carry predicates); a "+ carry * EPS" correction is a SEL (mask from
// the carry predicate) and a 64-bit add, or with GL_WIDE_CORR one IMAD.WIDE.U32 (carry * 0xffffffff + t);
// x2 * EPS + (x1:x0) is an IADD3 / IMAD.HI pair with carry-out.  Never feed an add-chain carry into subc (ptxas keeps the hardware not-borrow convention
// across the mix); add chains end in addc, sub chains in subc.
// On the host the same names return canonical values (a valid lazy representative).
#if defined(__CUDACC__)
__device__ __forceinline__ void mul_limbs(uint64_t a, uint64_t b, uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
    unsigned __int128 p = (unsigned __int128)a * b;
    uint64_t lo = (uint64_t)p, hi = (uint64_t)(p >> 64);
    x0 = (uint32_t)lo;
    x1 = (uint32_t)(lo >> 32);
    x2 = (uint32_t)hi;
    x3 = (uint32_t)(hi >> 32);
}
#endif

// GL_WIDE_CORR = 1 applies every "+ carry * EPS" correction as one IMAD.WIDE (fewest instructions); 0 (default) as a
// masked 64-bit add on the ALU pipe (one instruction more, but IMAD.WIDE occupies the fmaheavy pipe for 4 cycles and
// that pipe is the binding resource of the Poseidon and NTT kernels: profiles/r01m_*).
#ifndef GL_WIDE_CORR
#define GL_WIDE_CORR 0
#endif
#if defined(__CUDACC__)
// t + (c ? EPS : 0) for a carry c in {0, 1}; the caller guarantees the sum does not wrap
__device__ __forceinline__ uint64_t plus_carry_eps(uint64_t t, uint32_t c) {
#if GL_WIDE_CORR
    return t + (uint64_t)c * 0xFFFFFFFFu;
#else
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 m, t0, t1; .reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "selp.b32    m, 0xffffffff, 0, p;\n\t"
        "mov.b64     {t0, t1}, %1;\n\t"
        "add.cc.u32  t0, t0, m;\n\t"
        "addc.u32    t1, t1, 0;\n\t"
        "mov.b64     %0, {t0, t1};\n\t"
        "}"
        : "=l"(r)
        : "l"(t), "r"(c));
    return r;
#endif
}
#endif

GL_HD uint64_t canon_fast(uint64_t a) {
#if defined(__CUDA_ARCH__)
#if GL_WIDE_CORR
    uint32_t c;
    asm("{\n\t"
        ".reg .u32 a0, a1, s0, s1;\n\t"
        "mov.b64     {a0, a1}, %1;\n\t"
        "add.cc.u32  s0, a0, 0xffffffff;\n\t"  // a + EPS wraps  <=>  a >= p
        "addc.cc.u32 s1, a1, 0;\n\t"
        "addc.u32    %0, 0, 0;\n\t"
        "}"
        : "=r"(c)
        : "l"(a));
    return a + (uint64_t)c * 0xFFFFFFFFu;  // a - p == a + EPS (mod 2^64)
#else
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 a0, a1, s0, s1, c; .reg .pred p;\n\t"
        "mov.b64     {a0, a1}, %1;\n\t"
        "add.cc.u32  s0, a0, 0xffffffff;\n\t"  // s = a + EPS = a - p (mod 2^64); wraps  <=>  a >= p
        "addc.cc.u32 s1, a1, 0;\n\t"
        "addc.u32    c, 0, 0;\n\t"
        "setp.ne.u32 p, c, 0;\n\t"
        "selp.b32    s0, s0, a0, p;\n\t"
        "selp.b32    s1, s1, a1, p;\n\t"
        "mov.b64     %0, {s0, s1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a));
    return r;
#endif
#else
    return canon(a);
#endif
}
GL_HD uint64_t add_lc(uint64_t a, uint64_t v) {
#if defined(__CUDA_ARCH__)
    uint64_t t;
    uint32_t c;
    asm("add.cc.u64  %0, %2, %3;\n\t"
        "addc.u32    %1, 0, 0;"
        : "=l"(t), "=r"(c)
        : "l"(a), "l"(v));
    return plus_carry_eps(t, c);  // t = a + v - 2^64 < v < p: t + EPS cannot wrap again
#else
    return add(canon(a), v);
#endif
}
GL_HD uint64_t sub_lc(uint64_t a, uint64_t v) {
#if defined(__CUDA_ARCH__)
    uint64_t t;
    uint32_t m;
    asm("sub.cc.u64  %0, %2, %3;\n\t"
        "subc.u32    %1, 0, 0;"  // m = borrow ? 0xffffffff : 0
        : "=l"(t), "=r"(m)
        : "l"(a), "l"(v));
    return t - (uint64_t)m;  // t = a - v + 2^64 > EPS: t - EPS cannot underflow again
#else
    return sub(canon(a), v);
#endif
}
#if defined(__CUDACC__)
// x = x0 + x1*2^32 + x2*2^64 + x3*2^96  ==  (x1:x0) + x2*(2^32 - 1) - x3   (mod p)
__device__ __forceinline__ uint64_t reduce_limbs(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    uint32_t t0, t1, c;
    asm("mad.lo.cc.u32   %0, %5, 0xffffffff, %3;\n\t"  // (t1:t0) = x2 * EPS + (x1:x0), carry out
        "madc.hi.cc.u32  %1, %5, 0xffffffff, %4;\n\t"
        "addc.u32        %2, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(c)
        : "r"(x0), "r"(x1), "r"(x2));
    // wrapped value < 2^64 - 2^33, so + EPS cannot wrap again
    const uint64_t t = plus_carry_eps(((uint64_t)t1 << 32) | t0, c);
    return sub_lc(t, (uint64_t)x3);
}
#endif
GL_HD uint64_t mul_lazy(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t x0, x1, x2, x3;
    mul_limbs(a, b, x0, x1, x2, x3);
    return reduce_limbs(x0, x1, x2, x3);
#else
    uint64_t lo, hi;
    mul_wide(a, b, lo, hi);
    return reduce128(lo, hi);
#endif
}
// a^2 with three IMAD.WIDE instead of four (a0*a1 is used twice): one fmaheavy slot less per squaring at the price of
// three ALU instructions -- for the kernels whose binding pipe is fmaheavy (the Poseidon S-boxes)
GL_HD uint64_t sqr_lazy(uint64_t a) {
#if defined(__CUDA_ARCH__)
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    const uint64_t p0 = (uint64_t)a0 * a0, p1 = (uint64_t)a0 * a1, p2 = (uint64_t)a1 * a1;
    const unsigned __int128 v = (unsigned __int128)p0 + ((unsigned __int128)p1 << 33) + ((unsigned __int128)p2 << 64);
    const uint64_t lo = (uint64_t)v, hi = (uint64_t)(v >> 64);
    return reduce_limbs((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
#else
    return mul_lazy(a, a);
#endif
}
GL_HD uint64_t mad_lazy(uint64_t a, uint64_t b, uint64_t c) {
#if defined(__CUDA_ARCH__)
    unsigned __int128 p = (unsigned __int128)a * b + c;  // < 2^128: (2^64-1)^2 + 2^64 - 1
    uint64_t lo = (uint64_t)p, hi = (uint64_t)(p >> 64);
    return reduce_limbs((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
#else
    uint64_t lo, hi;
    mul_wide(a, b, lo, hi);
    return add(reduce128(lo, hi), canon(c));
#endif
}
// lo + x2 * 2^64 with a 32-bit x2 (sums of small-constant products: the MDS layer)
GL_HD uint64_t reduce96_lazy(uint64_t lo, uint32_t x2) {
#if defined(__CUDA_ARCH__)
    uint32_t t0, t1, c;
    asm("mad.lo.cc.u32   %0, %5, 0xffffffff, %3;\n\t"
        "madc.hi.cc.u32  %1, %5, 0xffffffff, %4;\n\t"
        "addc.u32        %2, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(c)
        : "r"((uint32_t)lo), "r"((uint32_t)(lo >> 32)), "r"(x2));
    return plus_carry_eps(((uint64_t)t1 << 32) | t0, c);
#else
    return reduce128(lo, (uint64_t)x2);
#endif
}
GL_HD uint64_t reduce128_lazy(uint64_t lo, uint64_t hi) {
#if defined(__CUDA_ARCH__)
    return reduce_limbs((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
#else
    return reduce128(lo, hi);
#endif
}
GL_HD uint64_t mul(uint64_t a, uint64_t b) { return canon_fast(mul_lazy(a, b)); }
GL_HD uint64_t sqr(uint64_t a) { return mul(a, a); }
// a*b + c  (c canonical)
GL_HD uint64_t mad(uint64_t a, uint64_t b, uint64_t c) { return add(mul(a, b), c); }

GL_HD uint64_t pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = sqr(b);
        e >>= 1;
    }
    return r;
}
GL_HD uint64_t inv(uint64_t a) { return pow(a, P - 2); }

// types.rs:240-244 primitive_root_of_unity
GL_HD uint64_t root_of_unity(int n_log) {
    uint64_t b = TWO_ADIC_GEN;
    for (int i = 0; i < TWO_ADICITY - n_log; i++) b = sqr(b);
    return b;
}

// ---- quadratic extension F[X]/(X^2 - 7) ----
struct ext2 {
    uint64_t c0, c1;
};
GL_HD ext2 make2(uint64_t a, uint64_t b) {
    ext2 r;
    r.c0 = a;
    r.c1 = b;
    return r;
}
GL_HD ext2 add(ext2 a, ext2 b) { return make2(add(a.c0, b.c0), add(a.c1, b.c1)); }
GL_HD ext2 sub(ext2 a, ext2 b) { return make2(sub(a.c0, b.c0), sub(a.c1, b.c1)); }
GL_HD ext2 neg(ext2 a) { return make2(neg(a.c0), neg(a.c1)); }
GL_HD uint64_t mul7(uint64_t a) {
    // 7a = 8a - a, three doublings and a subtraction: cheaper than a full product
    uint64_t a2 = dbl(a), a4 = dbl(a2), a8 = dbl(a4);
    return sub(a8, a);
}
GL_HD ext2 mul(ext2 a, ext2 b) {
    uint64_t c0 = add(mul(a.c0, b.c0), mul7(mul(a.c1, b.c1)));
    uint64_t c1 = add(mul(a.c0, b.c1), mul(a.c1, b.c0));
    return make2(c0, c1);
}
GL_HD ext2 mul(ext2 a, uint64_t s) { return make2(mul(a.c0, s), mul(a.c1, s)); }
GL_HD ext2 sqr(ext2 a) { return mul(a, a); }
GL_HD bool eq(ext2 a, ext2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }
GL_HD ext2 inv(ext2 a) {
    uint64_t n = sub(sqr(a.c0), mul7(sqr(a.c1)));
    uint64_t ni = inv(n);
    return make2(mul(a.c0, ni), mul(neg(a.c1), ni));
}
GL_HD ext2 pow(ext2 b, uint64_t e) {
    ext2 r = make2(1, 0);
    while (e) {
        if (e & 1) r = mul(r, b);
        b = sqr(b);
        e >>= 1;
    }
    return r;
}

GL_HD uint32_t bitrev32(uint32_t x, int bits) {
    if (bits == 0) return 0;
#if defined(__CUDA_ARCH__)
    return __brev(x) >> (32 - bits);
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

}  // namespace gl
