// SASS instruction counts of the Goldilocks primitives in olavm_b200/csrc/gl.cuh (no GPU needed):
//   for op in 0 1 2 3 4 5; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DOP=$op -I olavm_b200/csrc -cubin \
//       -o /tmp/prims_$op.cubin tools/microbench/prims.cu; cuobjdump -sass /tmp/prims_$op.cubin | grep -cE '^\s+/\*[0-9a-f]{4}\*/'; done
// (count(op) - count(5)) / 8 = instructions per primitive (ops 0, 2, 3, 4 include ~1.5 for loading the second operand).
//   0 mul_lazy   1 canon_fast   2 add_lc   3 sub_lc   4 one Cooley-Tukey butterfly   5 empty (baseline)
#include "gl.cuh"
#define N 8
extern "C" __global__ void k(uint64_t* x, const uint64_t* w) {
    int i = threadIdx.x;
    uint64_t a[N], b[N], t[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        a[j] = x[i + 64 * j];
        b[j] = x[i + 64 * j + 1024];
        t[j] = w[i + 64 * j];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
#if OP == 0
        a[j] = gl::mul_lazy(b[j], t[j]);
#elif OP == 1
        a[j] = gl::canon_fast(b[j]);
#elif OP == 2
        a[j] = gl::add_lc(b[j], t[j]);
#elif OP == 3
        a[j] = gl::sub_lc(b[j], t[j]);
#elif OP == 4
        {
            uint64_t v = gl::canon_fast(gl::mul_lazy(b[j], t[j]));
            uint64_t u = a[j];
            a[j] = gl::add_lc(u, v);
            b[j] = gl::sub_lc(u, v);
        }
#else
        a[j] = b[j];
#endif
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        x[i + 64 * j] = a[j];
        x[i + 64 * j + 1024] = b[j];
    }
}
