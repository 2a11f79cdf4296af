// 96-bit lazy Goldilocks arithmetic for the register rounds of the forward NTT network (ntt_tile.cuh).
//
// With t = 2^32 the Goldilocks prime is p = t^2 - t + 1, so t is a primitive 6th root of unity mod p:
//     t^2 = t - 1,   t^3 = -1   (mod p),
// and 2 has order 192: every 2^k-th root of unity with k <= 6 is a power of two (omega_64 = 2^39 for the reference's
// generator, goldilocks_field.rs:77).  A radix-16 block of the Cooley-Tukey network (four stages on sixteen rows,
// bfly_regs<4> in ntt_tile.cuh) evaluates  sum_m x_m z^m  at the sixteen points z = theta * omega_16^j: scaling the
// inputs by theta^m (fifteen general products, one per element instead of two) leaves a plain 16-point transform whose
// twiddles are all powers of two.  Inside that transform a value is a 96-bit two's-complement integer
//     V = a + b t + c t^2      (a, b unsigned words, c a signed word with a few significant bits)
// congruent to the field element: additions and subtractions are three-instruction carry chains with NO modular
// correction, a multiplication by 2^(32 q + r) is a funnel shift by r followed by one fold with the identities above,
// and only the sixteen outputs are brought back to 64 bits.  Nothing here is floating point or approximate: every
// step is an identity mod p, and the result differs from the canonical-form butterflies only in the representative
// ("lazy" u64), which the last pass canonicalises as before.
//
// Replaces (as one step of the same network) the arithmetic of cfft's butterflies, plonky2/field/src/cfft/serial.rs.
#pragma once
#include "gl.cuh"

namespace gl {

struct W96 {
    uint32_t a, b;
    int32_t c;
};

GL_HD W96 w96_from_u64(uint64_t x) {
    W96 r;
    r.a = (uint32_t)x;
    r.b = (uint32_t)(x >> 32);
    r.c = 0;
    return r;
}

// assemble  A + B t  from two small signed sums (|A|, |B| < 2^40)
GL_HD W96 w96_from_ab(int64_t A, int64_t B) {
    W96 r;
    r.a = (uint32_t)A;
    const int64_t U = B + (A >> 32);
    r.b = (uint32_t)U;
    r.c = (int32_t)(U >> 32);
    return r;
}

GL_HD W96 w96_add(W96 x, W96 y) {
    W96 r;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %3, %6;\n\t"
        "addc.cc.u32 %1, %4, %7;\n\t"
        "addc.u32 %2, %5, %8;"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c)
        : "r"(x.a), "r"(x.b), "r"(x.c), "r"(y.a), "r"(y.b), "r"(y.c));
#else
    const uint64_t s0 = (uint64_t)x.a + y.a;
    const uint64_t s1 = (uint64_t)x.b + y.b + (s0 >> 32);
    r.a = (uint32_t)s0;
    r.b = (uint32_t)s1;
    r.c = (int32_t)((uint32_t)x.c + (uint32_t)y.c + (uint32_t)(s1 >> 32));
#endif
    return r;
}
GL_HD W96 w96_sub(W96 x, W96 y) {
    W96 r;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %3, %6;\n\t"
        "subc.cc.u32 %1, %4, %7;\n\t"
        "subc.u32 %2, %5, %8;"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c)
        : "r"(x.a), "r"(x.b), "r"(x.c), "r"(y.a), "r"(y.b), "r"(y.c));
#else
    const uint64_t s0 = (uint64_t)x.a - y.a;
    const uint64_t s1 = (uint64_t)x.b - y.b - ((s0 >> 32) & 1);
    r.a = (uint32_t)s0;
    r.b = (uint32_t)s1;
    r.c = (int32_t)((uint32_t)x.c - (uint32_t)y.c - (uint32_t)((s1 >> 32) & 1));
#endif
    return r;
}

#if defined(__CUDA_ARCH__)
// x0 + x1 t + x2 t^2 - x3  for four unsigned words, t^2 = t - 1 folding x2:
//   (t1:t0, carry) = x2 * (2^32 - 1) + (x1:x0),  then x3 is subtracted; the top word ends up in {-1, 0, 1}
__device__ __forceinline__ W96 w96_fold(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    W96 r;
    asm("{\n\t"
        ".reg .u32 t0, t1, t2;\n\t"
        "mad.lo.cc.u32  t0, %5, 0xffffffff, %3;\n\t"
        "madc.hi.cc.u32 t1, %5, 0xffffffff, %4;\n\t"
        "addc.u32       t2, 0, 0;\n\t"
        "sub.cc.u32     %0, t0, %6;\n\t"
        "subc.cc.u32    %1, t1, 0;\n\t"
        "subc.u32       %2, t2, 0;\n\t"
        "}"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
    return r;
}
#endif

// x * w for two u64 (any representatives): the 128-bit product x0 + x1 t + x2 t^2 + x3 t^3 folded with
// t^2 = t - 1, t^3 = -1:  (x0 - x2 - x3) + (x1 + x2) t
GL_HD W96 w96_mul(uint64_t x, uint64_t w) {
#if defined(__CUDA_ARCH__)
    uint32_t x0, x1, x2, x3;
    mul_limbs(x, w, x0, x1, x2, x3);
    return w96_fold(x0, x1, x2, x3);
#else
    const unsigned __int128 p = (unsigned __int128)x * w;
    const uint32_t x0 = (uint32_t)p, x1 = (uint32_t)(p >> 32), x2 = (uint32_t)(p >> 64), x3 = (uint32_t)(p >> 96);
    return w96_from_ab((int64_t)x0 - (int64_t)x2 - (int64_t)x3, (int64_t)x1 + (int64_t)x2);
#endif
}

// V * 2^E for a compile-time 0 < E < 96 (E = 32 q + r, 0 < r < 32) and -2^(32-r) <= c < 2^(32-r) (the register rounds
// keep |c| <= 15 at every shifted operand, ntt_shift.cuh).  The 128-bit shift V << r is w0 + w1 t + w2 t^2 + w3 t^3 with
// w3 = the sign word of c (0 or -1) under that bound, and w3 t^3 = -w3 = n := [c < 0]: the low r bits of w0 = a << r are
// zero, so w0' = (a << r) + n absorbs it without a carry and
//     V 2^r = w0' + w1 t + w2 t^2   (mod p),  all three words unsigned.
// Times t^q, folded with t^2 = t - 1, t^3 = -1:
//   q = 0:  (w1:w0') + w2 (2^32 - 1)
//   q = 1:  (w0' + w1) t - (w1 + w2)  =  w1 (2^32 - 1) + (w0' t - w2)
//   q = 2:  (w0' - w2) t - (w0' + w1) =  w0' (2^32 - 1) - (w1 + w2 t)
template <int E>
GL_HD W96 w96_mul_pow2(W96 v) {
    static_assert(E > 0 && E < 96 && E % 32 != 0, "shift out of range");
    constexpr int q = E / 32, r = E % 32;
    const uint32_t n = (uint32_t)v.c >> 31;
    const uint32_t w0 = (v.a << r) + n;
#if defined(__CUDA_ARCH__)
    const uint32_t w1 = __funnelshift_l(v.a, v.b, r);
    const uint32_t w2 = __funnelshift_l(v.b, (uint32_t)v.c, r);
    W96 o;
    if (q == 0) {
        asm("mad.lo.cc.u32  %0, %5, 0xffffffff, %3;\n\t"
            "madc.hi.cc.u32 %1, %5, 0xffffffff, %4;\n\t"
            "addc.u32       %2, 0, 0;"
            : "=r"(o.a), "=r"(o.b), "=r"(o.c)
            : "r"(w0), "r"(w1), "r"(w2));
        return o;
    }
    // q = 1:  w1 (t - 1) + (w0' t - w2);   q = 2:  w0' (t - 1) - (w1 + w2 t):  a negated two-word value, then one multiply-add
    // chain by 2^32 - 1 (IADD3 + IMAD.HI with carry) into it
    const uint32_t m = (q == 1) ? w1 : w0;
    if (q == 1) {
        asm("sub.cc.u32  %0, 0, %4;\n\t"
            "subc.cc.u32 %1, %3, 0;\n\t"
            "subc.u32    %2, 0, 0;"
            : "=r"(o.a), "=r"(o.b), "=r"(o.c)
            : "r"(w0), "r"(w2));
    } else {
        asm("sub.cc.u32  %0, 0, %3;\n\t"
            "subc.cc.u32 %1, 0, %4;\n\t"
            "subc.u32    %2, 0, 0;"
            : "=r"(o.a), "=r"(o.b), "=r"(o.c)
            : "r"(w1), "r"(w2));
    }
    asm("mad.lo.cc.u32  %0, %3, 0xffffffff, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, 0xffffffff, %1;\n\t"
        "addc.u32       %2, %2, 0;"
        : "+r"(o.a), "+r"(o.b), "+r"(o.c)
        : "r"(m));
    return o;
#else
#if defined(W96_CHECK_BOUNDS)
    if (v.c < -(1 << (32 - r)) || v.c >= (1 << (32 - r))) {
        fprintf(stderr, "w96_mul_pow2<%d>: top word %d outside the bound the shift needs\n", E, v.c);
        abort();
    }
#endif
    const uint32_t w1 = (v.b << r) | (v.a >> (32 - r));
    const uint32_t w2 = ((uint32_t)v.c << r) | (v.b >> (32 - r));
    const int64_t W0 = (int64_t)(uint64_t)w0, W1 = (int64_t)(uint64_t)w1, W2 = (int64_t)(uint64_t)w2;
    if (q == 0) return w96_from_ab(W0 - W2, W1 + W2);
    if (q == 1) return w96_from_ab(-W1 - W2, W0 + W1);
    return w96_from_ab(-W0 - W1, W0 - W2);
#endif
}

// A multiple of p added to ONE input of a block whose every output contains that input with coefficient 1 (row 0 of
// the 2^K-point transform) makes every output non-negative: 2^12 p > 2^75 bounds the magnitude of any sum of sixteen
// folded terms (each below 2^67), and w96_to_u64_nonneg then needs no sign handling.
GL_HD W96 w96_bias(W96 v) {
    // 2^12 p = 2^76 - 2^44 + 2^12:  words (0x1000, 0xfffff000, 0xfff)
    W96 k;
    k.a = 0x1000u;
    k.b = 0xfffff000u;
    k.c = 0xfff;
    return w96_add(v, k);
}

// back to a u64 representative (any value in [0, 2^64) congruent to V) for 0 <= c < 2^30:
//   a + b t + c t^2 = (b:a) + c (2^32 - 1): one wrap at most, folded as + (2^32 - 1); the wrapped value is below
//   2^62, so the correction cannot wrap again
GL_HD uint64_t w96_to_u64_nonneg(W96 v) {
#if !defined(__CUDA_ARCH__) && defined(W96_CHECK_BOUNDS)
    if (v.c < 0 || v.c >= (1 << 30)) {
        fprintf(stderr, "w96_to_u64_nonneg: top word %d\n", v.c);
        abort();
    }
#endif
#if defined(__CUDA_ARCH__)
    uint32_t t0, t1, k;
    asm("mad.lo.cc.u32  %0, %5, 0xffffffff, %3;\n\t"
        "madc.hi.cc.u32 %1, %5, 0xffffffff, %4;\n\t"
        "addc.u32       %2, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(k)
        : "r"(v.a), "r"(v.b), "r"((uint32_t)v.c));
    return plus_carry_eps(((uint64_t)t1 << 32) | t0, k);
#else
    const unsigned __int128 s = (unsigned __int128)(((uint64_t)v.b << 32) | v.a) + (unsigned __int128)(uint32_t)v.c * EPS;
    return (uint64_t)s + (uint64_t)(s >> 64) * EPS;
#endif
}

// the general version (signed c, |c| < 2^30):
//   a + b t + c (t - 1) = (a - c) + (b + c) t  = lo + u0 t + u1 t^2 with u1 in {-1, 0, 1};  the second fold cannot wrap.
GL_HD uint64_t w96_to_u64(W96 v) {
    const int64_t A = (int64_t)(uint64_t)v.a - (int64_t)v.c;
    const int64_t U = (int64_t)(uint64_t)v.b + (int64_t)v.c + (A >> 32);
    const uint64_t x = ((uint64_t)(uint32_t)U << 32) | (uint32_t)A;
    const uint64_t u1 = (uint64_t)(U >> 32);  // -1, 0 or 1
    return x + (u1 << 32) - u1;
}

}  // namespace gl
