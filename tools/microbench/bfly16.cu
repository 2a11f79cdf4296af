// Register-round butterflies in isolation: the canonical-product radix-2 form (bfly_regs of ntt_tile.cuh) against the
// shift-twiddle form (ntt_shift.cuh), same 2^K-row block, same twiddle semantics.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I olavm_b200/csrc -o /tmp/bfly16 tools/microbench/bfly16.cu
//   /tmp/bfly16 host        CPU check: shift form == reference butterflies (canonical values), K = 1..4
//   /tmp/bfly16             GPU: equality of the two forms + cycles per warp-butterfly at several occupancies
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#if !defined(__CUDACC__)
#define W96_CHECK_BOUNDS 1  // the host build aborts if a shifted operand or an output leaves the range the device code assumes
#endif
#include "gl.cuh"
#include "ntt_shift.cuh"

using namespace ola::ntt::tile;

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

// twiddle of (stage s, sub-block ql) of a block whose deepest twiddle is theta
static uint64_t old_twiddle(int K, int s, int ql, uint64_t theta, bool inv = false) {
    uint64_t th = theta;
    for (int i = 0; i < K - 1 - s; ++i) th = gl::sqr(th);
    uint64_t w = gl::root_of_unity(s + 1);
    if (inv) w = gl::inv(w);
    int br = 0;
    for (int i = 0; i < s; ++i) br |= ((ql >> i) & 1) << (s - 1 - i);
    return gl::mul(th, gl::pow(w, (uint64_t)br));
}

template <int K, bool INV = false>
static int host_check(int trials) {
    constexpr int NE = 1 << K;
    int bad = 0;
    for (int tr = 0; tr < trials; ++tr) {
        uint64_t theta = gl::canon(rnd());
        if (tr == 0) theta = 1;
        if (tr == 1) theta = gl::P - 1;
        if (tr % 11 == 2) theta = ~0ULL - (rnd() & 1);  // a lazy twiddle is never stored, but the product accepts it
        uint64_t x[NE], ref[NE], v[NE][1], tw[NE];
        for (int m = 0; m < NE; ++m) {
            x[m] = rnd();
            if (tr % 7 == 3) x[m] = ~0ULL - (rnd() & 3);  // non-canonical extremes
            if (tr % 7 == 5) x[m] = (rnd() & 1) ? gl::P - 1 : 0;
            if (tr % 7 == 6) x[m] = (rnd() & 1) ? ~0ULL : ((rnd() & 1) ? 0xFFFFFFFF00000000ULL : 0x00000000FFFFFFFFULL);
            ref[m] = gl::canon(x[m]);
            v[m][0] = x[m];
        }
        // reference: the radix-2 stages with canonical arithmetic
        for (int s = 0; s < K; ++s) {
            const int half = (NE >> 1) >> s;
            for (int ql = 0; ql < (1 << s); ++ql) {
                const uint64_t w = old_twiddle(K, s, ql, theta, INV);
                for (int jj = 0; jj < half; ++jj) {
                    const int m = (ql << (K - s)) + jj;
                    const uint64_t p = gl::mul(ref[m + half], w), a = ref[m];
                    ref[m] = gl::add(a, p);
                    ref[m + half] = gl::sub(a, p);
                }
            }
        }
        // every third trial folds an output scale into the table (the iNTT's 1/n): tw = theta^m * scale, row 0 scaled explicitly
        const bool scaled = (tr % 3 == 1);
        const uint64_t scale = gl::canon(rnd());
        uint64_t pw = scaled ? scale : 1;
        for (int m = 1; m < NE; ++m) {
            pw = gl::mul(pw, theta);
            tw[m - 1] = pw;
        }
        if (scaled)
            for (int m = 0; m < NE; ++m) ref[m] = gl::mul(ref[m], scale);
        bfly_shift<K, 0, 1, INV>(v, tw, 0, scaled ? &scale : nullptr);
        for (int m = 0; m < NE; ++m)
            if (gl::canon(v[m][0]) != ref[m]) {
                if (bad < 5) printf("K=%d trial %d row %d: got %016llx want %016llx\n", K, tr, m, (unsigned long long)gl::canon(v[m][0]), (unsigned long long)ref[m]);
                ++bad;
            }
    }
    printf("host check K=%d%s: %s (%d trials)\n", K, INV ? " (inverse roots)" : "", bad ? "FAILED" : "ok", trials);
    return bad;
}

#if defined(__CUDACC__)
// the shipped radix-2 form (ntt_tile.cuh bfly_regs, forward direction)
template <int K, int U0, int LN>
__device__ __forceinline__ void bfly_old(uint64_t (&v)[1 << K][LN], const uint64_t* __restrict__ t) {
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const int half = (1 << (K - 1)) >> s;
#pragma unroll
        for (int ql = 0; ql < (1 << s); ++ql) {
            const uint64_t w = t[(((1 << s) - 1) + ql) << U0];
#pragma unroll
            for (int jj = 0; jj < half; ++jj) {
                const int m = (ql << (K - s)) + jj;
#pragma unroll
                for (int ln = 0; ln < LN; ++ln) {
                    const uint64_t p = gl::canon_fast(gl::mul_lazy(v[m + half][ln], w));
                    const uint64_t a = v[m][ln];
                    v[m][ln] = gl::add_lc(a, p);
                    v[m + half][ln] = gl::sub_lc(a, p);
                }
            }
        }
    }
}

// FORM 0 = old, 1 = shift.  tw: [64 blocks][2^K - 1] per form; every iteration uses another block's twiddles (LDS)
template <int K, int LN, int FORM>
__global__ void __launch_bounds__(256) bench_kernel(uint64_t* data, const uint64_t* tw_g, int iters) {
    constexpr int NE = 1 << K;
    __shared__ uint64_t tw[64 * (NE - 1)];
    for (int i = threadIdx.x; i < 64 * (NE - 1); i += blockDim.x) tw[i] = tw_g[i];
    __syncthreads();
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t v[NE][LN];
#pragma unroll
    for (int m = 0; m < NE; ++m)
#pragma unroll
        for (int ln = 0; ln < LN; ++ln) v[m][ln] = data[(gid * NE + m) * LN + ln];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const uint64_t* t = tw + ((it + threadIdx.x / 32) & 63) * (NE - 1);
        if (FORM == 0)
            bfly_old<K, 0, LN>(v, t);
        else
            bfly_shift<K, 0, LN>(v, t);
    }
#pragma unroll
    for (int m = 0; m < NE; ++m)
#pragma unroll
        for (int ln = 0; ln < LN; ++ln) data[(gid * NE + m) * LN + ln] = gl::canon_fast(v[m][ln]);
}

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

template <int K, int LN>
static int gpu_run(int blocks_per_sm, int iters) {
    constexpr int NE = 1 << K;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int threads = 256, blocks = sms * blocks_per_sm;
    const size_t n = (size_t)blocks * threads * NE * LN;
    std::vector<uint64_t> h(n), tw_old(64 * (NE - 1)), tw_new(64 * (NE - 1));
    for (auto& x : h) x = rnd();
    for (int b = 0; b < 64; ++b) {
        const uint64_t theta = gl::canon(rnd());
        for (int s = 0; s < K; ++s)
            for (int ql = 0; ql < (1 << s); ++ql) tw_old[b * (NE - 1) + ((1 << s) - 1) + ql] = old_twiddle(K, s, ql, theta);
        uint64_t pw = 1;
        for (int m = 1; m < NE; ++m) {
            pw = gl::mul(pw, theta);
            tw_new[b * (NE - 1) + m - 1] = pw;
        }
    }
    uint64_t *d0, *d1, *t0, *t1;
    CK(cudaMalloc(&d0, n * 8));
    CK(cudaMalloc(&d1, n * 8));
    CK(cudaMalloc(&t0, tw_old.size() * 8));
    CK(cudaMalloc(&t1, tw_new.size() * 8));
    CK(cudaMemcpy(t0, tw_old.data(), tw_old.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t1, tw_new.data(), tw_new.size() * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms[2] = {0, 0};
    for (int form = 0; form < 2; ++form) {
        uint64_t* d = form ? d1 : d0;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e0));
            if (form == 0)
                bench_kernel<K, LN, 0><<<blocks, threads>>>(d, t0, iters);
            else
                bench_kernel<K, LN, 1><<<blocks, threads>>>(d, t1, iters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (rep == 0 || t < ms[form]) ms[form] = t;
        }
    }
    std::vector<uint64_t> r0(n), r1(n);
    CK(cudaMemcpy(r0.data(), d0, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r1.data(), d1, n * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad += (r0[i] != r1[i]);
    const double clk = prop.clockRate * 1e3;  // Hz (nominal; the box runs at its boost clock, see bench.py clocks)
    const double warps = (double)blocks * threads / 32.0;
    const double bf = warps * iters * (double)(K * NE / 2) * LN;  // warp-butterflies
    for (int form = 0; form < 2; ++form)
        printf("K=%d LN=%d CTAs/SM=%d (%d warps/SMSP) %-5s %8.3f ms  %6.2f cycles per warp-butterfly per SMSP (at %.0f MHz nominal)\n", K, LN, blocks_per_sm,
               blocks_per_sm * threads / 32 / 4, form ? "shift" : "old", ms[form], ms[form] * 1e-3 * clk * sms * 4 / bf, clk / 1e6);
    printf("   equality of canonical outputs after %d chained rounds: %s (%zu of %zu differ)\n", iters, bad ? "FAILED" : "ok", bad, n);
    cudaFree(d0);
    cudaFree(d1);
    cudaFree(t0);
    cudaFree(t1);
    return bad != 0;
}
#endif

int main(int argc, char** argv) {
    int bad = 0;
    bad += host_check<1>(2000);
    bad += host_check<2>(2000);
    bad += host_check<3>(2000);
    bad += host_check<4>(4000);
    bad += host_check<2, true>(2000);
    bad += host_check<3, true>(2000);
    bad += host_check<4, true>(4000);
    // the constant the shift form rests on
    if (gl::root_of_unity(6) != gl::pow(2, 39)) {
        printf("omega_64 != 2^39\n");
        ++bad;
    }
    if (argc > 1 && !strcmp(argv[1], "host")) return bad != 0;
#if defined(__CUDACC__)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        printf("no GPU: host checks only\n");
        return bad != 0;
    }
    const int iters = 2000;
    for (int bps : {1, 2, 4}) {
        bad += gpu_run<4, 1>(bps, iters);
        bad += gpu_run<3, 2>(bps, iters);
    }
    bad += gpu_run<4, 2>(2, iters);
    bad += gpu_run<3, 1>(4, iters);
#endif
    printf(bad ? "FAILED\n" : "PASSED\n");
    return bad != 0;
}
