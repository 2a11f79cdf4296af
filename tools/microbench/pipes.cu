// Integer-pipe throughput micro-benchmark for sm_100a: how many warp-instructions per cycle per SM sub-partition do
// IMAD.WIDE.U32, IMAD (lo), IMAD.HI, IADD3(.X), LOP3, SHF and mixes of them sustain?  The Goldilocks kernels (NTT,
// Poseidon, quotient) are bound by these pipes, not by HBM; the numbers calibrate the per-kernel instruction budgets
// in DESIGN.md.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

// 8 independent chains per thread, each op repeated; kinds selected at compile time.
template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, long long* cyc) {
    uint32_t a[8], b[8], c[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * (threadIdx.x + 1 + i);
        b[i] = a[i] ^ 0x9e3779b9u;
        c[i] = a[i] + 77u * i;
        d[i] = b[i] + 13u;
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // every operation consumes its own previous result, so nothing is loop-invariant
            if (KIND == 0) {  // IMAD.WIDE.U32 with 64-bit accumulate: (b:a) = a * c + (b:a)
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0,%1}, t;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));
            } else if (KIND == 1) {  // IMAD lo
                asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(c[i]));
            } else if (KIND == 2) {  // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c[i]), "r"(d[i]));
            } else if (KIND == 3) {  // IADD3 + IADD3.X pair (64-bit add)
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(d[i]));
            } else if (KIND == 4) {  // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c[i]), "r"(d[i]));
            } else if (KIND == 5) {  // SHF
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(c[i]));
            } else if (KIND == 6) {  // 1 IMAD.WIDE : 2 IADD3
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0,%1}, t;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c[i]), "+r"(d[i]) : "r"(a[i]), "r"(b[i]));
            } else if (KIND == 7) {  // 1 IMAD.WIDE : 4 IADD3
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0,%1}, t;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c[i]), "+r"(d[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c[i]), "+r"(d[i]) : "r"(b[i]), "r"(a[i]));
            } else if (KIND == 8) {  // 1 IMAD : 1 IADD3
                asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(c[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
            } else if (KIND == 9) {  // 1 IMAD.WIDE : 1 IMAD : 2 IADD3
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0,%1}, t;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(d[i]));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c[i]), "+r"(d[i]) : "r"(a[i]), "r"(b[i]));
            } else if (KIND == 10) {  // IMUL lo (IMAD with RZ addend)
                asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c[i]));
            } else if (KIND == 11) {  // IMAD.WIDE.U32 without accumulate: (b:a) = a * c
                asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0,%1}, t;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));
            } else if (KIND == 12) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(a[i]) : "r"(c[i]));
            } else if (KIND == 13) {  // ISETP + SEL
                asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a[i]) : "r"(c[i]), "r"(d[i]));
            } else if (KIND == 14) {  // IMAD.X-style: add with carry-in only
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, 0;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));
            } else if (KIND == 15) {  // MOV chain (register rotation)
                asm volatile("{.reg .u32 t; mov.u32 t, %0; mov.u32 %0, %1; mov.u32 %1, %2; mov.u32 %2, t;}" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i] ^ b[i] ^ c[i] ^ d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_inner, uint32_t* out, long long* cyc, int nsm) {
    k<KIND><<<nsm, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<nsm, 1024>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // 32 warps per SM = 8 per SMSP; ptxas rewrites the PTX (splits accumulates, converts adds to IMAD.X ...), so the
    // instruction mix of each loop body is read from cuobjdump (tools/microbench/pipes_report.py) and combined with the
    // cycles per loop iteration printed here
    (void)ops_per_inner;
    printf("KIND %2d %-40s cycles/iteration(8 warps per SMSP)=%8.2f  (%.3f ms)\n", KIND, name, (double)h[0] / ITER, ms);
}

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)nsm * 1024 * 4);
    cudaMalloc(&cyc, nsm * 8);
    printf("SMs=%d\n", nsm);
    run<0>("IMAD.WIDE.U32 (acc)", 1, out, cyc, nsm);
    run<11>("IMAD.WIDE.U32 (no acc)", 1, out, cyc, nsm);
    run<1>("IMAD lo", 1, out, cyc, nsm);
    run<10>("IMUL lo", 1, out, cyc, nsm);
    run<2>("IMAD.HI", 1, out, cyc, nsm);
    run<3>("IADD3 + IADD3.X", 2, out, cyc, nsm);
    run<4>("LOP3", 1, out, cyc, nsm);
    run<5>("SHF", 1, out, cyc, nsm);
    run<12>("PRMT", 1, out, cyc, nsm);
    run<13>("ISETP+SEL", 2, out, cyc, nsm);
    run<6>("1 IMAD.WIDE : 2 IADD3", 3, out, cyc, nsm);
    run<7>("1 IMAD.WIDE : 4 IADD3", 5, out, cyc, nsm);
    run<8>("1 IMAD : 1 IADD3", 2, out, cyc, nsm);
    run<9>("1 IMAD.WIDE : 1 IMAD : 2 IADD3", 4, out, cyc, nsm);
    run<14>("IADD3 + carry-only add", 2, out, cyc, nsm);
    run<15>("3 MOV (rotation)", 3, out, cyc, nsm);
    return 0;
}
