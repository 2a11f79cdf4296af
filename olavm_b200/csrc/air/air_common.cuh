// Device-side vocabulary for AIR constraint evaluation (the B200 counterpart of `PackedField` arithmetic,
// `StarkEvaluationVars` and `ConstraintConsumer`: circuits/src/stark/{vars.rs:5-13, constraint_consumer.rs:10-80}).
// One thread evaluates one LDE row (P::WIDTH == 1 semantics).
#pragma once
#include "../gl.cuh"
#include "cpu_air.h"
#include "mem_air.h"

namespace ola {
namespace air {

struct Fl;  // an unreduced ("lazy") product
struct Fp {
    uint64_t v;
    __host__ __device__ __forceinline__ Fp() : v(0) {}
    __host__ __device__ __forceinline__ explicit Fp(uint64_t x) : v(x) {}
    __host__ __device__ __forceinline__ Fp(const Fl& l);  // canonicalises
    __host__ __device__ __forceinline__ Fp& operator+=(Fp o);
    __host__ __device__ __forceinline__ Fp& operator*=(Fp o);
};
// A product that has been reduced to 64 bits but not canonicalised: any representative in [0, 2^64).  Products feed
// further products and the constraint consumer as they are (both accept any representative); only an addition or a
// subtraction needs the canonical value, and the implicit conversion to Fp supplies it (4 instructions, paid only there).
struct Fl {
    uint64_t v;
    __host__ __device__ __forceinline__ explicit Fl(uint64_t x) : v(x) {}
};
__host__ __device__ __forceinline__ Fp::Fp(const Fl& l) : v(gl::canon_fast(l.v)) {}
// canonical in, canonical out.  a - b is the borrow-corrected difference (5 instructions); a + b = a - (p - b)
// (7; p - 0 = p is not canonical but is a valid subtrahend: a - p borrows and the correction gives a back)
__host__ __device__ __forceinline__ Fp operator+(Fp a, Fp b) { return Fp(gl::sub_lc(a.v, gl::P - b.v)); }
__host__ __device__ __forceinline__ Fp operator-(Fp a, Fp b) { return Fp(gl::sub_lc(a.v, b.v)); }
__host__ __device__ __forceinline__ Fl operator*(Fp a, Fp b) { return Fl(gl::mul_lazy(a.v, b.v)); }
__host__ __device__ __forceinline__ Fl operator*(Fl a, Fp b) { return Fl(gl::mul_lazy(a.v, b.v)); }
__host__ __device__ __forceinline__ Fl operator*(Fp a, Fl b) { return Fl(gl::mul_lazy(a.v, b.v)); }
__host__ __device__ __forceinline__ Fl operator*(Fl a, Fl b) { return Fl(gl::mul_lazy(a.v, b.v)); }
__host__ __device__ __forceinline__ Fp operator+(Fl a, Fp b) { return Fp(a) + b; }
__host__ __device__ __forceinline__ Fp operator+(Fp a, Fl b) { return a + Fp(b); }
__host__ __device__ __forceinline__ Fp operator+(Fl a, Fl b) { return Fp(a) + Fp(b); }
__host__ __device__ __forceinline__ Fp operator-(Fl a, Fp b) { return Fp(a) - b; }
__host__ __device__ __forceinline__ Fp operator-(Fp a, Fl b) { return a - Fp(b); }
__host__ __device__ __forceinline__ Fp operator-(Fl a, Fl b) { return Fp(a) - Fp(b); }
__host__ __device__ __forceinline__ Fp& Fp::operator+=(Fp o) {
    v = gl::sub_lc(v, gl::P - o.v);
    return *this;
}
__host__ __device__ __forceinline__ Fp& Fp::operator*=(Fp o) {
    v = gl::mul(v, o.v);
    return *this;
}
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
    // each column's low 64 bits live in one u64 so that ptxas keeps them in an aligned register pair, which is what
    // IMAD.WIDE accumulates into (separate u32 halves cost two IMAD.MOV per multiply-accumulate to re-pair them)
    uint64_t c0, c1, c2;
    uint32_t t0, t1, t2;
    __device__ __forceinline__ void clear() {
        c0 = c1 = c2 = 0;
        t0 = t1 = t2 = 0;
    }
    __device__ __forceinline__ void mac(uint64_t a, uint64_t b) {
        asm("{\n\t"
            ".reg .u32 a0, a1, b0, b1, l, h;\n\t"
            "mov.b64 {a0, a1}, %6;\n\t"
            "mov.b64 {b0, b1}, %7;\n\t"
            "mov.b64 {l, h}, %0;\n\t"
            "mad.lo.cc.u32  l, a0, b0, l;\n\t"
            "madc.hi.cc.u32 h, a0, b0, h;\n\t"
            "addc.u32       %3, %3, 0;\n\t"
            "mov.b64 %0, {l, h};\n\t"
            "mov.b64 {l, h}, %1;\n\t"
            "mad.lo.cc.u32  l, a0, b1, l;\n\t"
            "madc.hi.cc.u32 h, a0, b1, h;\n\t"
            "addc.u32       %4, %4, 0;\n\t"
            "mad.lo.cc.u32  l, a1, b0, l;\n\t"
            "madc.hi.cc.u32 h, a1, b0, h;\n\t"
            "addc.u32       %4, %4, 0;\n\t"
            "mov.b64 %1, {l, h};\n\t"
            "mov.b64 {l, h}, %2;\n\t"
            "mad.lo.cc.u32  l, a1, b1, l;\n\t"
            "madc.hi.cc.u32 h, a1, b1, h;\n\t"
            "addc.u32       %5, %5, 0;\n\t"
            "mov.b64 %2, {l, h};\n\t"
            "}"
            : "+l"(c0), "+l"(c1), "+l"(c2), "+r"(t0), "+r"(t1), "+r"(t2)
            : "l"(a), "l"(b));
    }
    // canonical value of the sum: x = C0 + C1 2^32 + C2 2^64 as five 32-bit limbs; 2^128 = -2^32 (mod p)
    __device__ __forceinline__ uint64_t reduce() const {
        const uint32_t c0l = (uint32_t)c0, c0h = (uint32_t)(c0 >> 32), c1l = (uint32_t)c1, c1h = (uint32_t)(c1 >> 32), c2l = (uint32_t)c2,
                       c2h = (uint32_t)(c2 >> 32);
        uint32_t x1, x2, x3, x4;
        asm("add.cc.u32  %0, %4, %6;\n\t"   // x1 = c0h + c1l
            "addc.cc.u32 %1, %5, %7;\n\t"   // x2 = t0 + c1h
            "addc.cc.u32 %2, %8, 0;\n\t"    // x3 = t1
            "addc.u32    %3, 0, 0;\n\t"     // x4
            "add.cc.u32  %1, %1, %9;\n\t"   // x2 += c2l
            "addc.cc.u32 %2, %2, %10;\n\t"  // x3 += c2h
            "addc.u32    %3, %3, %11;"       // x4 += t2
            : "=&r"(x1), "=&r"(x2), "=&r"(x3), "=&r"(x4)
            : "r"(c0h), "r"(t0), "r"(c1l), "r"(c1h), "r"(t1), "r"(c2l), "r"(c2h), "r"(t2));
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
    Wide acc0, acc1;  // sum of weighted constraints that are not transition constraints
    Wide tr0, tr1;    // sum of weighted transition constraints: multiplied by z_last once, at the end
    const uint64_t* w;
    int k;
    Fp z_last, lagrange_first, lagrange_last;
    __device__ __forceinline__ void add_weighted(Wide& a0, Wide& a1, uint64_t c) {
        const ulonglong2 ww = *reinterpret_cast<const ulonglong2*>(w + 2 * k);
        a0.mac(c, ww.x);
        a1.mac(c, ww.y);
        ++k;
    }
    __device__ __forceinline__ void constraint(Fp c) { add_weighted(acc0, acc1, c.v); }
    __device__ __forceinline__ void constraint(Fl c) { add_weighted(acc0, acc1, c.v); }
    __device__ __forceinline__ void constraint_transition(Fp c) { add_weighted(tr0, tr1, c.v); }
    __device__ __forceinline__ void constraint_transition(Fl c) { add_weighted(tr0, tr1, c.v); }
    __device__ __forceinline__ void constraint_first_row(Fp c) { constraint(c * lagrange_first); }
    __device__ __forceinline__ void constraint_first_row(Fl c) { constraint(c * lagrange_first); }
    __device__ __forceinline__ void constraint_last_row(Fp c) { constraint(c * lagrange_last); }
    __device__ __forceinline__ void constraint_last_row(Fl c) { constraint(c * lagrange_last); }
    // the two accumulated values sum_k c_k alpha_j^(K-1-k), canonical
    __device__ __forceinline__ void finish(Fp& out0, Fp& out1) const {
        out0 = Fp(acc0.reduce()) + Fp(tr0.reduce()) * z_last;
        out1 = Fp(acc1.reduce()) + Fp(tr1.reduce()) * z_last;
    }
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
