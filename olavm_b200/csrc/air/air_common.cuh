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
    __host__ __device__ __forceinline__ Fp operator+(Fp o) const { return Fp(gl::add(v, o.v)); }
    __host__ __device__ __forceinline__ Fp operator-(Fp o) const { return Fp(gl::sub(v, o.v)); }
    __host__ __device__ __forceinline__ Fp operator*(Fp o) const { return Fp(gl::mul(v, o.v)); }
    __host__ __device__ __forceinline__ Fp& operator+=(Fp o) {
        v = gl::add(v, o.v);
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

// ConstraintConsumer with num_challenges == 2 (constraint_consumer.rs:46-80)
struct Consumer {
    Fp alpha0, alpha1, acc0, acc1, z_last, lagrange_first, lagrange_last;
    __host__ __device__ __forceinline__ void constraint(Fp c) {
        acc0 = acc0 * alpha0 + c;
        acc1 = acc1 * alpha1 + c;
    }
    __host__ __device__ __forceinline__ void constraint_transition(Fp c) { constraint(c * z_last); }
    __host__ __device__ __forceinline__ void constraint_first_row(Fp c) { constraint(c * lagrange_first); }
    __host__ __device__ __forceinline__ void constraint_last_row(Fp c) { constraint(c * lagrange_last); }
};

// circuits/src/stark/lookup.rs:13-35
__host__ __device__ __forceinline__ void eval_lookups(const Row& lv, const Row& nv, Consumer& yc, int col_permuted_input, int col_permuted_table) {
    Fp local_perm_input = lv[col_permuted_input];
    Fp next_perm_table = nv[col_permuted_table];
    Fp next_perm_input = nv[col_permuted_input];
    Fp diff_input_prev = next_perm_input - local_perm_input;
    Fp diff_input_table = next_perm_input - next_perm_table;
    yc.constraint(diff_input_prev * diff_input_table);
    yc.constraint_last_row(diff_input_table);
}

}  // namespace air
}  // namespace ola
