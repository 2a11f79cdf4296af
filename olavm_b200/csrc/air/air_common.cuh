// Device-side vocabulary for AIR constraint evaluation (the B200 counterpart of `PackedField` arithmetic,
// `StarkEvaluationVars` and `ConstraintConsumer`: circuits/src/stark/{vars.rs:5-13, constraint_consumer.rs:10-80}).
// One thread evaluates one LDE row (P::WIDTH == 1 semantics).
#pragma once
#include "../gl.cuh"
#include "cpu_air.h"
#include "mem_air.h"

namespace ola {
namespace air {

struct Fp {
    uint64_t v;
    __host__ __device__ __forceinline__ Fp() : v(0) {}
    __host__ __device__ __forceinline__ explicit Fp(uint64_t x) : v(x) {}
    // canonical in, canonical out.  a - b is the borrow-corrected difference (5 instructions); a + b = a - (p - b)
    // (7; p - 0 = p is not canonical but is a valid subtrahend: a - p borrows and the correction gives a back)
    __host__ __device__ __forceinline__ Fp operator+(Fp o) const { return Fp(gl::sub_lc(v, gl::P - o.v)); }
    __host__ __device__ __forceinline__ Fp operator-(Fp o) const { return Fp(gl::sub_lc(v, o.v)); }
    __host__ __device__ __forceinline__ Fp operator*(Fp o) const { return Fp(gl::mul(v, o.v)); }
    __host__ __device__ __forceinline__ Fp& operator+=(Fp o) {
        v = gl::sub_lc(v, gl::P - o.v);
        return *this;
    }
    __host__ __device__ __forceinline__ Fp& operator*=(Fp o) {
        v = gl::mul(v, o.v);
        return *this;
    }
};
__host__ __device__ __forceinline__ Fp fp(uint64_t k) { return Fp(k); }  // k must be canonical
template <>
__host__ __device__ __forceinline__ Fp kc<Fp>(uint64_t k) {
    return Fp(k);
}
template <>
__host__ __device__ __forceinline__ bool is_zero<Fp>(const Fp& x) {
    return x.v == 0;
}
__host__ __device__ __forceinline__ Fp one() { return Fp(1); }

// one LDE row of a column-major batch: element c at base[c*stride + r]
struct Row {
    const uint64_t* __restrict__ base;
    size_t stride, r;
    __host__ __device__ __forceinline__ Fp operator[](int c) const {
#if defined(__CUDA_ARCH__)
        return Fp(__ldg(base + (size_t)c * stride + r));
#else
        return Fp(base[(size_t)c * stride + r]);
#endif
    }
};

#if defined(__CUDACC__)
// An unreduced sum of 64 x 64-bit products: three 96-bit columns (a0*b0 | a0*b1 + a1*b0 | a1*b1, weights 1, 2^32, 2^64).
// One multiply-accumulate is 4 IMAD.WIDE.U32 with carry-out + 2 IADD3.X -- against 21 (multiply, canonicalise) + 7 (add)
// for the same step in reduced arithmetic.  Operands are arbitrary u64 representatives; up to 2^31 terms.
struct Wide {
    uint32_t c0l, c0h, c0t, c1l, c1h, c1t, c2l, c2h, c2t;
    __device__ __forceinline__ void clear() { c0l = c0h = c0t = c1l = c1h = c1t = c2l = c2h = c2t = 0; }
    __device__ __forceinline__ void mac(uint64_t a, uint64_t b) {
        const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
        asm("mad.lo.cc.u32  %0, %9, %11, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %11, %1;\n\t"
            "addc.u32       %2, %2, 0;\n\t"
            "mad.lo.cc.u32  %3, %9, %12, %3;\n\t"
            "madc.hi.cc.u32 %4, %9, %12, %4;\n\t"
            "addc.u32       %5, %5, 0;\n\t"
            "mad.lo.cc.u32  %3, %10, %11, %3;\n\t"
            "madc.hi.cc.u32 %4, %10, %11, %4;\n\t"
            "addc.u32       %5, %5, 0;\n\t"
            "mad.lo.cc.u32  %6, %10, %12, %6;\n\t"
            "madc.hi.cc.u32 %7, %10, %12, %7;\n\t"
            "addc.u32       %8, %8, 0;"
            : "+r"(c0l), "+r"(c0h), "+r"(c0t), "+r"(c1l), "+r"(c1h), "+r"(c1t), "+r"(c2l), "+r"(c2h), "+r"(c2t)
            : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    }
    // canonical value of the sum: x = C0 + C1 2^32 + C2 2^64 as five 32-bit limbs; 2^128 = -2^32 (mod p)
    __device__ __forceinline__ uint64_t reduce() const {
        uint32_t x1, x2, x3, x4;
        asm("add.cc.u32  %0, %4, %6;\n\t"   // x1 = c0h + c1l
            "addc.cc.u32 %1, %5, %7;\n\t"   // x2 = c0t + c1h
            "addc.cc.u32 %2, %8, 0;\n\t"    // x3 = c1t
            "addc.u32    %3, 0, 0;\n\t"     // x4
            "add.cc.u32  %1, %1, %9;\n\t"   // x2 += c2l
            "addc.cc.u32 %2, %2, %10;\n\t"  // x3 += c2h
            "addc.u32    %3, %3, %11;"       // x4 += c2t
            : "=&r"(x1), "=&r"(x2), "=&r"(x3), "=&r"(x4)
            : "r"(c0h), "r"(c0t), "r"(c1l), "r"(c1h), "r"(c1t), "r"(c2l), "r"(c2h), "r"(c2t));
        const uint64_t lo = gl::canon_fast(gl::reduce_limbs(c0l, x1, x2, x3));
        return gl::sub_lc(lo, (uint64_t)x4 << 32);  // x4 2^32 <= 2^64 - 2^32 < p
    }
};

// ConstraintConsumer with num_challenges == 2 (constraint_consumer.rs:46-80).  The reference accumulates by Horner's rule,
// acc_j = acc_j alpha_j + c; the same field element is sum_k c_k alpha_j^(K-1-k) over the table's K constraints, which
// needs no reduction between terms: weights w[2k + j] = alpha_j^(K-1-k) are read from shared memory (one LDS.128 per
// constraint) and each constraint costs two Wide::mac.  k counts the constraints emitted so far (a compile-time constant
// wherever the AIR's loops are unrolled).
struct Consumer {
    Wide acc0, acc1;
    const uint64_t* w;
    int k;
    Fp z_last, lagrange_first, lagrange_last;
    __device__ __forceinline__ void constraint(Fp c) {
        const ulonglong2 ww = *reinterpret_cast<const ulonglong2*>(w + 2 * k);
        acc0.mac(c.v, ww.x);
        acc1.mac(c.v, ww.y);
        ++k;
    }
    __device__ __forceinline__ void constraint_transition(Fp c) { constraint(c * z_last); }
    __device__ __forceinline__ void constraint_first_row(Fp c) { constraint(c * lagrange_first); }
    __device__ __forceinline__ void constraint_last_row(Fp c) { constraint(c * lagrange_last); }
};

// circuits/src/stark/lookup.rs:13-35
__device__ __forceinline__ void eval_lookups(const Row& lv, const Row& nv, Consumer& yc, int col_permuted_input, int col_permuted_table) {
    Fp local_perm_input = lv[col_permuted_input];
    Fp next_perm_table = nv[col_permuted_table];
    Fp next_perm_input = nv[col_permuted_input];
    Fp diff_input_prev = next_perm_input - local_perm_input;
    Fp diff_input_table = next_perm_input - next_perm_table;
    yc.constraint(diff_input_prev * diff_input_table);
    yc.constraint_last_row(diff_input_table);
}
#endif  // __CUDACC__

}  // namespace air
}  // namespace ola
