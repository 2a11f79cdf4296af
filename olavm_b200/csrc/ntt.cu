// Goldilocks NTT / iNTT / coset-LDE kernels for sm_100a.
//
// Replaces, on the proving hot path:
//   plonky2/field/src/cfft/{mod,serial,concurrent}.rs   evaluate_poly, interpolate_poly,
//       evaluate_poly_with_offset (coset LDE), interpolate_poly_with_offset (coset iNTT)
//   plonky2/plonky2/src/fri/oracle.rs:84-85             transpose + reverse_index_bits_in_place
//       (fused away: the un-permuted output of the forward network IS the Merkle leaf order)
//
// Design (B200-first, not a translation of the reference's recursion):
//   * One butterfly network, two directions.  FORWARD = Cooley-Tukey butterflies (multiply, then
//     add/sub) over natural-order input, producing bit-reversed *positions*; INVERSE = Gentleman-Sande
//     butterflies undoing it.  A coset shift s is folded into the twiddles (stage-u twiddle of the
//     sub-block evaluating on sigma*H_m is sigma^(m/2^(u+1)) * omega_{2^(u+1)}^{bitrev(q)}), so a coset
//     LDE costs exactly the butterflies of a plain NTT: no separate "clone_and_shift" pass
//     (cfft/concurrent.rs:186) and no inter-pass four-step twiddle multiply.
//   * A transform of 2^L points is cut into passes of <= 11 stages.  Each pass stages a tile in
//     shared memory: the first pass(es) take tiles of T=8 adjacent low indices x 2^l strided rows
//     (64-byte coalesced segments), the last pass takes contiguous 2^l-element runs.  HBM traffic is
//     one read + one write of the data per pass.
//   * Twiddles: per tile, 2^l - 1 products  c_u * BRS[q]  built in shared memory from a 2048-entry
//     bit-reversed root table (L1/L2 resident) and l per-row constants c_u obtained from a 3-level
//     power table of omega_{2^32} (5120 entries).  No O(n) twiddle table is ever streamed from HBM.
//   * Column batches, cosets and tiles map to blockIdx.{y,z,x}: thousands of CTAs per launch, a
//     multiple-wave fill of the 148 SMs for every real trace shape.
//   * This is 64-bit modular integer arithmetic: IMAD/IADD3 pipes only, no tensor-core formulation.
#include "common.h"
#include "gl.cuh"
#include "ntt.h"

#include <cstdlib>

namespace ola {
namespace ntt {

static constexpr int MAX_PASS_LOG = 11;
static constexpr int TILE_T = 8;

struct PassArgs {
    const uint64_t* src;
    uint64_t* dst;
    size_t src_col_stride, dst_col_stride;      // elements between consecutive columns
    size_t src_coset_stride, dst_coset_stride;  // elements between consecutive cosets (blockIdx.z)
    const uint64_t* pw;   // 3-level omega_{2^32} power tables (direction chosen by the host)
    const uint64_t* brs;  // bit-reversed small root table (same direction)
    int L;                // log2 of the transform size
    int M;                // log2 of the sub-block size entering this pass
    int l;                // stages handled by this pass
    int G;                // contiguous pass: sub-blocks per CTA
    int bitrev_store;     // contiguous pass: write to bit-reversed (natural-order) addresses
    size_t out_mul;       // bitrev_store: address = natural_index*out_mul + bitrev(coset)  (interleaved cosets)
    int coset_bits;
    int apply_scale;
    size_t tiles_per_cta;  // tiled passes: tiles (sharing one twiddle table) streamed through each CTA
    int lazy_out;         // tile kernels: leave the outputs un-canonicalised (an intermediate pass; every pass accepts lazy input)
    size_t ncols;         // tile_contig: columns in the batch (the last column group may be partial)
    uint64_t scale;
    int prefetch;         // contiguous tile passes: bulk-prefetch the tile this many groups ahead into L2 (0 = off)
    int inv_roots;        // pw / brs hold the inverse roots (the shift form of the tile kernels needs to know which powers of two they are)
    uint64_t s_last[MAX_COSETS];  // per coset: (shift_i)^(2^(M-l)), or its inverse for the GS network
};

__device__ __forceinline__ uint64_t pow_omega(const uint64_t* __restrict__ pw, uint32_t E) {
    uint64_t r = __ldg(pw + 4096 + (E >> 22));
    r = gl::mul(r, __ldg(pw + 2048 + ((E >> 11) & 2047u)));
    r = gl::mul(r, __ldg(pw + (E & 2047u)));
    return r;
}

// c[u] for u = 0..l-1 of the sub-block Q (t = L - M stages already done):
//   c[l-1] = s_last * omega_{2^(t+l)}^{bitrev_t(Q)},  c[u-1] = c[u]^2.
__device__ __forceinline__ void row_constants(const PassArgs& a, uint64_t s_last, uint32_t Q, uint64_t* c) {
    const int t = a.L - a.M;
    uint32_t e = gl::bitrev32(Q, t);
    uint32_t E = (t == 0) ? 0u : (e << (32 - t - a.l));
    uint64_t v = gl::mul(s_last, pow_omega(a.pw, E));
    for (int u = a.l - 1; u >= 0; --u) {
        c[u] = v;
        v = gl::sqr(v);
    }
}

// ---- strided pass: tile = (sub-block Q, T adjacent inner indices) x all 2^l rows -----------------
template <bool GS>
__global__ void __launch_bounds__(1024) pass_strided(const PassArgs a) {
    extern __shared__ uint64_t sm[];
    const int l = a.l, R = 1 << l, T = TILE_T;
    uint64_t* tw = sm;        // [R]   tw[2^u + q]
    uint64_t* cu = sm + R;    // [16]
    uint64_t* x = sm + R + 16;  // [R][T]
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t coset = blockIdx.z;
    const size_t inner = (size_t)1 << (a.M - l);
    const uint32_t tiles_per_sub = (uint32_t)(inner / T);
    const uint32_t Q = blockIdx.x / tiles_per_sub;
    const size_t c0 = (size_t)(blockIdx.x % tiles_per_sub) * T;
    const uint64_t* in = a.src + blockIdx.y * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << a.M) + c0;
    uint64_t* out = a.dst + blockIdx.y * a.dst_col_stride + coset * a.dst_coset_stride + ((size_t)Q << a.M) + c0;

    if (tid == 0) row_constants(a, a.s_last[coset], Q, cu);
    for (int i = tid; i < R * T; i += nt) {
        int r = i / T, c = i % T;
        x[i] = gl::canon(in[(size_t)r * inner + c]);
    }
    __syncthreads();
    for (int i = tid + 1; i < R; i += nt) {
        int u = 31 - __clz(i);
        tw[i] = gl::mul(cu[u], __ldg(a.brs + (i - (1 << u))));
    }
    __syncthreads();

    const int nb = (R / 2) * T;
    if (!GS) {
        for (int u = 0; u < l; ++u) {
            const int hs = l - 1 - u, half = 1 << hs;
            for (int b = tid; b < nb; b += nt) {
                int c = b % T, bb = b / T;
                int q = bb >> hs, j = bb & (half - 1);
                int i0 = ((q << (hs + 1)) + j) * T + c, i1 = i0 + half * T;
                uint64_t v = gl::mul(x[i1], tw[(1 << u) + q]);
                uint64_t w = x[i0];
                x[i0] = gl::add(w, v);
                x[i1] = gl::sub(w, v);
            }
            __syncthreads();
        }
    } else {
        for (int u = l - 1; u >= 0; --u) {
            const int hs = l - 1 - u, half = 1 << hs;
            for (int b = tid; b < nb; b += nt) {
                int c = b % T, bb = b / T;
                int q = bb >> hs, j = bb & (half - 1);
                int i0 = ((q << (hs + 1)) + j) * T + c, i1 = i0 + half * T;
                uint64_t A = x[i0], B = x[i1];
                x[i0] = gl::add(A, B);
                x[i1] = gl::mul(gl::sub(A, B), tw[(1 << u) + q]);
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < R * T; i += nt) {
        int r = i / T, c = i % T;
        uint64_t v = x[i];
        if (a.apply_scale) v = gl::mul(v, a.scale);
        out[(size_t)r * inner + c] = v;
    }
}

// ---- contiguous pass: G sub-blocks of 2^l consecutive elements (M == l) --------------------------
template <bool GS>
__global__ void __launch_bounds__(1024) pass_contig(const PassArgs a) {
    extern __shared__ uint64_t sm[];
    const int l = a.l, R = 1 << l, G = a.G;
    uint64_t* tw = sm;                 // [G][R]
    uint64_t* x = sm + (size_t)G * R;  // [G][R]
    uint64_t* cu = x + (size_t)G * R;  // [G][16]
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t coset = blockIdx.z;
    const int t = a.L - a.M;
    const uint32_t k0 = blockIdx.x * G;  // first sub-block (or first natural row index when bitrev_store)
    const uint64_t* in = a.src + blockIdx.y * a.src_col_stride + coset * a.src_coset_stride;
    uint64_t* out = a.dst + blockIdx.y * a.dst_col_stride + coset * a.dst_coset_stride;

    auto subblock = [&](int g) -> uint32_t { return a.bitrev_store ? gl::bitrev32(k0 + g, t) : (k0 + g); };

    if (tid < G) row_constants(a, a.s_last[coset], subblock(tid), cu + tid * 16);
    for (int i = tid; i < G * R; i += nt) {
        int g = i >> l, r = i & (R - 1);
        x[i] = gl::canon(in[((size_t)subblock(g) << l) + r]);
    }
    __syncthreads();
    for (int i = tid; i < G * R; i += nt) {
        int g = i >> l, k = i & (R - 1);
        if (k) {
            int u = 31 - __clz(k);
            tw[i] = gl::mul(cu[g * 16 + u], __ldg(a.brs + (k - (1 << u))));
        }
    }
    __syncthreads();

    const int nb = G * (R / 2);
    if (!GS) {
        for (int u = 0; u < l; ++u) {
            const int hs = l - 1 - u, half = 1 << hs;
            for (int b = tid; b < nb; b += nt) {
                int g = b >> (l - 1), bb = b & (R / 2 - 1);
                int q = bb >> hs, j = bb & (half - 1);
                int i0 = (g << l) + (q << (hs + 1)) + j, i1 = i0 + half;
                uint64_t v = gl::mul(x[i1], tw[(g << l) + (1 << u) + q]);
                uint64_t w = x[i0];
                x[i0] = gl::add(w, v);
                x[i1] = gl::sub(w, v);
            }
            __syncthreads();
        }
    } else {
        for (int u = l - 1; u >= 0; --u) {
            const int hs = l - 1 - u, half = 1 << hs;
            for (int b = tid; b < nb; b += nt) {
                int g = b >> (l - 1), bb = b & (R / 2 - 1);
                int q = bb >> hs, j = bb & (half - 1);
                int i0 = (g << l) + (q << (hs + 1)) + j, i1 = i0 + half;
                uint64_t A = x[i0], B = x[i1];
                x[i0] = gl::add(A, B);
                x[i1] = gl::mul(gl::sub(A, B), tw[(g << l) + (1 << u) + q]);
            }
            __syncthreads();
        }
    }
    if (!a.bitrev_store) {
        for (int i = tid; i < G * R; i += nt) {
            int g = i >> l, r = i & (R - 1);
            uint64_t v = x[i];
            if (a.apply_scale) v = gl::mul(v, a.scale);
            out[((size_t)(k0 + g) << l) + r] = v;
        }
    } else {
        // in-place position p = Q*R + r holds the value of natural index bitrev_L(p) = bitrev_l(r)*2^t + bitrev_t(Q)
        // = kk*2^t + (k0+g).  Consecutive g -> consecutive addresses (G*8-byte segments).
        // With cosets interleaved (natural-order LDE): address = (kk*2^t + k0+g)*out_mul + bitrev(coset).
        out = a.dst + blockIdx.y * a.dst_col_stride;
        const size_t add = a.coset_bits ? gl::bitrev32(coset, a.coset_bits) : 0;
        for (int i = tid; i < G * R; i += nt) {
            int g = i % G, kk = i / G;
            int r = gl::bitrev32((uint32_t)kk, l);
            uint64_t v = x[(g << l) + r];
            if (a.apply_scale) v = gl::mul(v, a.scale);
            out[(((size_t)kk << t) + k0 + g) * a.out_mul + add] = v;
        }
    }
}

// =====================================================================================================
// Register-blocked passes (l >= 6): every thread owns 8 elements and runs three butterfly stages on them
// in registers (radix-8 rounds); shared memory is touched once per round instead of once per stage, the
// first round reads straight from HBM and the last round writes straight back.  Shared-memory indices are
// padded (one slot per 16) so that the power-of-two strides of every round are bank-conflict-free.
// Butterfly arithmetic is "lazy": data stays in [0, 2^64), only the twiddle product is canonicalised.
// =====================================================================================================
__device__ __forceinline__ int phys(int i) { return i + (i >> 4); }

// three (or 3 - SKIP) stages on 8 elements.  Stage s pairs m with m + (4 >> s); its twiddle is
// tw[2^(u0+s) + (qh << s) + (m >> (3 - s))]  (u0 = first stage of the round, qh = block index at stage u0).
template <bool GS, int SKIP>
__device__ __forceinline__ void bfly8(uint64_t (&x)[8], const uint64_t* __restrict__ tw, int u0, int qh) {
    if (!GS) {
#pragma unroll
        for (int s = SKIP; s < 3; ++s) {
            const int half = 4 >> s;
            const uint64_t* t = tw + (1 << (u0 + s)) + (qh << s);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (m & half) continue;
                uint64_t v = gl::canon_fast(gl::mul_lazy(x[m + half], t[m >> (3 - s)]));
                uint64_t a = x[m];
                x[m] = gl::add_lc(a, v);
                x[m + half] = gl::sub_lc(a, v);
            }
        }
    } else {
#pragma unroll
        for (int s = 2; s >= SKIP; --s) {
            const int half = 4 >> s;
            const uint64_t* t = tw + (1 << (u0 + s)) + (qh << s);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (m & half) continue;
                uint64_t b = gl::canon_fast(x[m + half]);
                uint64_t a = x[m];
                x[m] = gl::add_lc(a, b);
                x[m + half] = gl::mul_lazy(gl::sub_lc(a, b), t[m >> (3 - s)]);
            }
        }
    }
}

template <bool GS>
__device__ __forceinline__ void bfly8_dispatch(uint64_t (&x)[8], const uint64_t* tw, int u0, int qh, int skip) {
    if (skip == 0)
        bfly8<GS, 0>(x, tw, u0, qh);
    else if (skip == 1)
        bfly8<GS, 1>(x, tw, u0, qh);
    else
        bfly8<GS, 2>(x, tw, u0, qh);
}

// round schedule: full rounds at u0 = 0, 3, 6, ...; if l % 3 != 0 a final round at u0 = l - 3 that skips the
// stages already done.  CT runs rounds 0..nr-1, GS runs them backwards.
struct Rounds {
    int l, nfull, nr;
    __device__ Rounds(int l_) : l(l_), nfull(l_ / 3), nr(l_ / 3 + ((l_ % 3) ? 1 : 0)) {}
    __device__ int u0(int rho) const { return rho < nfull ? 3 * rho : l - 3; }
    __device__ int skip(int rho) const { return rho < nfull ? 0 : 3 - (l % 3); }
};

template <bool GS>
__global__ void __launch_bounds__(1024, 1) pass_strided_r8(const PassArgs a) {
    extern __shared__ uint64_t sm[];
    const int l = a.l, R = 1 << l;
    uint64_t* tw = sm;           // [R]
    uint64_t* cu = sm + R;       // [16]
    uint64_t* x = sm + R + 16;   // [phys(R * 8)]
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t coset = blockIdx.z;
    const size_t inner = (size_t)1 << (a.M - l);
    const uint32_t tiles_per_sub = (uint32_t)(inner / TILE_T);
    const uint32_t Q = blockIdx.x / tiles_per_sub;
    const size_t c0 = (size_t)(blockIdx.x % tiles_per_sub) * TILE_T;
    const uint64_t* in = a.src + blockIdx.y * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << a.M) + c0;
    uint64_t* out = a.dst + blockIdx.y * a.dst_col_stride + coset * a.dst_coset_stride + ((size_t)Q << a.M) + c0;

    if (tid == 0) row_constants(a, a.s_last[coset], Q, cu);
    __syncthreads();
    for (int i = tid + 1; i < R; i += nt) {
        int u = 31 - __clz(i);
        tw[i] = gl::mul(cu[u], __ldg(a.brs + (i - (1 << u))));
    }
    __syncthreads();

    const Rounds rd(l);
    for (int k = 0; k < rd.nr; ++k) {
        const int rho = GS ? rd.nr - 1 - k : k;
        const int u0 = rd.u0(rho), skip = rd.skip(rho);
        const bool first = (k == 0), last = (k == rd.nr - 1);
        const int sh = l - u0 - 3;
        for (int w = tid; w < R; w += nt) {
            const int c = w & 7, t = w >> 3;
            const int j = t & ((1 << sh) - 1), qh = t >> sh;
            const int rbase = (qh << (l - u0)) + j;
            uint64_t v[8];
            if (first) {
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = gl::canon_fast(in[(size_t)(rbase + (m << sh)) * inner + c]);
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = x[phys(((rbase + (m << sh)) << 3) + c)];
            }
            bfly8_dispatch<GS>(v, tw, u0, qh, skip);
            if (last) {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    uint64_t y = a.apply_scale ? gl::mul(v[m], a.scale) : gl::canon_fast(v[m]);
                    out[(size_t)(rbase + (m << sh)) * inner + c] = y;
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) x[phys(((rbase + (m << sh)) << 3) + c)] = v[m];
            }
        }
        if (!last) __syncthreads();
    }
}

template <bool GS>
__global__ void __launch_bounds__(1024) pass_contig_r8(const PassArgs a) {
    extern __shared__ uint64_t sm[];
    const int l = a.l, R = 1 << l, G = a.G;
    uint64_t* tw = sm;                      // [G][R]
    uint64_t* cu = sm + (size_t)G * R;      // [G][16]
    uint64_t* x = cu + (size_t)G * 16;      // [phys(G * R)]
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t coset = blockIdx.z;
    const int t_done = a.L - a.M;
    const uint32_t k0 = blockIdx.x * G;
    const uint64_t* in = a.src + blockIdx.y * a.src_col_stride + coset * a.src_coset_stride;
    uint64_t* out = a.dst + blockIdx.y * a.dst_col_stride + coset * a.dst_coset_stride;
    auto subblock = [&](int g) -> uint32_t { return a.bitrev_store ? gl::bitrev32(k0 + g, t_done) : (k0 + g); };

    if (tid < G) row_constants(a, a.s_last[coset], subblock(tid), cu + tid * 16);
    __syncthreads();
    for (int i = tid; i < G * R; i += nt) {
        int g = i >> l, k = i & (R - 1);
        if (k) {
            int u = 31 - __clz(k);
            tw[i] = gl::mul(cu[g * 16 + u], __ldg(a.brs + (k - (1 << u))));
        }
    }
    __syncthreads();

    const Rounds rd(l);
    const int items_per_g = R >> 3;
    for (int k = 0; k < rd.nr; ++k) {
        const int rho = GS ? rd.nr - 1 - k : k;
        const int u0 = rd.u0(rho), skip = rd.skip(rho);
        const bool first = (k == 0), last = (k == rd.nr - 1) && !a.bitrev_store;
        const int sh = l - u0 - 3;
        for (int w = tid; w < G * items_per_g; w += nt) {
            const int g = w / items_per_g, t = w - g * items_per_g;
            const int j = t & ((1 << sh) - 1), qh = t >> sh;
            const int rbase = (qh << (l - u0)) + j;
            uint64_t v[8];
            if (first) {
                const uint64_t* p = in + ((size_t)subblock(g) << l) + rbase;
                if (sh == 0) {
                    const ulonglong2* p2 = reinterpret_cast<const ulonglong2*>(p);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        ulonglong2 q = p2[m];
                        v[2 * m] = gl::canon_fast(q.x);
                        v[2 * m + 1] = gl::canon_fast(q.y);
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 8; ++m) v[m] = gl::canon_fast(p[(size_t)m << sh]);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = x[phys((g << l) + rbase + (m << sh))];
            }
            bfly8_dispatch<GS>(v, tw + ((size_t)g << l), u0, qh, skip);
            if (last) {
                uint64_t* p = out + ((size_t)(k0 + g) << l) + rbase;
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = a.apply_scale ? gl::mul(v[m], a.scale) : gl::canon_fast(v[m]);
                if (sh == 0) {
                    ulonglong2* p2 = reinterpret_cast<ulonglong2*>(p);
#pragma unroll
                    for (int m = 0; m < 4; ++m) p2[m] = make_ulonglong2(v[2 * m], v[2 * m + 1]);
                } else {
#pragma unroll
                    for (int m = 0; m < 8; ++m) p[(size_t)m << sh] = v[m];
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) x[phys((g << l) + rbase + (m << sh))] = v[m];
            }
        }
        if (!last) __syncthreads();
    }
    if (a.bitrev_store) {
        // position p = Q*R + r holds natural index bitrev_l(r)*2^t + (k0+g); consecutive g -> consecutive addresses
        uint64_t* o = a.dst + blockIdx.y * a.dst_col_stride;
        const size_t add = a.coset_bits ? gl::bitrev32(coset, a.coset_bits) : 0;
        for (int i = tid; i < G * R; i += nt) {
            int g = i % G, kk = i / G;
            int r = gl::bitrev32((uint32_t)kk, l);
            uint64_t y = x[phys((g << l) + r)];
            y = a.apply_scale ? gl::mul(y, a.scale) : gl::canon_fast(y);
            o[(((size_t)kk << t_done) + k0 + g) * a.out_mul + add] = y;
        }
    }
}

}  // namespace ntt
}  // namespace ola
#include "ntt_tile.cuh"
namespace ola {
namespace ntt {

// out[j] *= base * step^j   (coset un-shift after a natural-order inverse transform)
__global__ void scale_powers_kernel(uint64_t* data, size_t col_stride, size_t n, uint64_t base, uint64_t step) {
    const int RUN = 16;
    size_t j0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * RUN;
    if (j0 >= n) return;
    uint64_t* p = data + blockIdx.y * col_stride;
    uint64_t f = gl::mul(base, gl::pow(step, j0));
    for (int k = 0; k < RUN && j0 + k < n; ++k) {
        p[j0 + k] = gl::mul(gl::canon(p[j0 + k]), f);
        f = gl::mul(f, step);
    }
}

__global__ void build_tables_kernel(uint64_t* pw, uint64_t* brs, uint64_t omega /* omega_{2^32} or inverse */,
                                    uint64_t w12 /* omega_{2^12} or inverse */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2048) pw[i] = gl::pow(omega, (uint64_t)i);
    if (i < 2048) pw[2048 + i] = gl::pow(omega, (uint64_t)i << 11);
    if (i < 1024) pw[4096 + i] = gl::pow(omega, (uint64_t)i << 22);
    // BRS[i] = omega_{2^(k+1)}^{bitrev_k(i)} with k = 11: omega_{2^12}^{bitrev_11(i)}  (prefix property)
    if (i < 2048) brs[i] = gl::pow(w12, (uint64_t)gl::bitrev32((uint32_t)i, 11));
}

// ---- tiled passes (ntt_tile.cuh): configurations and dispatch -------------------------------------------------------
// OLA_NTT_VARIANT picks the thread mapping of the 10- and 11-stage passes (the LDE of 2^20 .. 2^22-row tables):
//   0: two lanes per thread in every round, 256 threads, 2 CTAs / SM
//   1: as 0 with one padding row per 32 and the register allocation capped for 3 CTAs / SM   (10-stage passes only)
//   2: two lanes in radix-8 rounds, one in radix-16 rounds: 512 threads of 16 elements, 2 CTAs / SM
//   3: one lane per thread: 512 threads, 2 CTAs / SM
static int tune_variant() {
    static int v = [] {
        const char* e = getenv("OLA_NTT_VARIANT");
        int t = e ? atoi(e) : 2;  // profiles/r01m_ntt_tile_sweeps.txt
        return (t >= 0 && t <= 3) ? t : 2;
    }();
    return v;
}
// OLA_NTT_SHIFT=0 runs the forward tile passes as radix-2 butterflies on canonical products (the round-1 form) instead
// of the shift-twiddle form of ntt_shift.cuh (A/B measurements: profiles/r02m_*)
static bool tune_shift() {
    static bool v = [] {
        const char* e = getenv("OLA_NTT_SHIFT");
        return !(e && atoi(e) == 0);
    }();
    return v;
}
// natural-order (bit-reversed store) last passes.  OLA_NTT_NAT = 2 (default): tile_nat, C sub-blocks per tile so that
// stores fill whole sectors; 0: the generic kernels of round 1 (16-byte pieces).  (One sub-block per CTA with
// single-element scattered stores was measured no better than the generic kernels and removed: profiles/r02n_*.)
static int tune_nat() {
    static int v = [] {
        const char* e = getenv("OLA_NTT_NAT");
        const int t = e ? atoi(e) : 2;
        return (t == 0) ? 0 : 2;
    }();
    return v;
}
// OLA_NTT_PREFETCH = tiles ahead whose columns the contiguous tile passes pull into L2 with cp.async.bulk.prefetch (0 = off)
static int tune_prefetch() {
    static int v = [] {
        const char* e = getenv("OLA_NTT_PREFETCH");
        const int t = e ? atoi(e) : 1;
        return (t >= 0 && t <= 4) ? t : 1;
    }();
    return v;
}
// OLA_NTT_STRIDED_C4=1: four-lane tiles (32-byte row segments, 256 threads, 4 CTAs / SM) in the strided 10-stage pass too
static bool tune_strided_c4() {
    static bool v = [] {
        const char* e = getenv("OLA_NTT_STRIDED_C4");
        return e && atoi(e) == 1;
    }();
    return v;
}
static bool tune_contig_c4() {
    static bool v = [] {
        const char* e = getenv("OLA_NTT_CONTIG_C4");
        return !(e && atoi(e) == 0);  // default on: 14.7 vs 15.7 ms on the 200 x 2^20 LDE (profiles/r01m_ntt_tile_sweeps.txt)
    }();
    return v;
}
template <typename G>
static void tile_optin(int max_optin) {
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_strided<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_strided<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
}
template <typename G>
static void tile_optin_shift(int max_optin, bool strided) {
    if (strided) {
        OLA_CUDA(cudaFuncSetAttribute(tile::tile_strided<G, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
        OLA_CUDA(cudaFuncSetAttribute(tile::tile_strided<G, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    }
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
}
template <typename G>
static void tile_optin_contig(int max_optin) {
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_contig<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
}
using T6 = tile::Cfg<6, 8>;
using T7 = tile::Cfg<7, 8>;
using T8 = tile::Cfg<8, 8>;
using T9 = tile::Cfg<9, 8>;
using T10 = tile::Cfg<10, 8>;
using T10v1 = tile::Cfg<10, 8, 2, 2, 5, 3>;
using T10v2 = tile::Cfg<10, 8, 2, 1, 4, 2>;
using T10v3 = tile::Cfg<10, 8, 1, 1, 4, 2>;
using T10c4 = tile::Cfg<10, 4, 2, 1, 4, 4>;  // contiguous pass, 4-column tiles: 43 KB, 4-5 CTAs / SM (OLA_NTT_CONTIG_C4=0 disables)
// (one lane per thread in the radix-8 rounds, and a 3-CTA register cap, were measured and dropped: profiles/r02p_*, r02s_*)
using T11 = tile::Cfg<11, 4>;
using T11v2 = tile::Cfg<11, 4, 2, 1, 4, 2>;
using T11v3 = tile::Cfg<11, 4, 1, 1, 4, 2>;
static void tile_optin_all(int max_optin) {
    tile_optin<T6>(max_optin);
    tile_optin<T7>(max_optin);
    tile_optin<T8>(max_optin);
    tile_optin<T9>(max_optin);
    tile_optin<T10>(max_optin);
    tile_optin<T10v1>(max_optin);
    tile_optin<T10v2>(max_optin);
    tile_optin<T10v3>(max_optin);
    tile_optin_contig<T10c4>(max_optin);

    tile_optin<T11>(max_optin);
    tile_optin<T11v2>(max_optin);
    tile_optin<T11v3>(max_optin);
    tile_optin_contig<tile::Cfg<6, 2>>(max_optin);
    tile_optin_contig<tile::Cfg<7, 2>>(max_optin);
    tile_optin_contig<tile::Cfg<8, 2>>(max_optin);
    tile_optin_contig<tile::Cfg<9, 2>>(max_optin);
    tile_optin_contig<tile::Cfg<10, 2>>(max_optin);
    tile_optin_contig<tile::Cfg<11, 2>>(max_optin);
    tile_optin_shift<T6>(max_optin, true);
    tile_optin_shift<T7>(max_optin, true);
    tile_optin_shift<T8>(max_optin, true);
    tile_optin_shift<T9>(max_optin, true);
    tile_optin_shift<T10v2>(max_optin, true);
    tile_optin_shift<T10c4>(max_optin, true);
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_strided<T10c4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_strided<T10c4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));

    tile_optin_shift<T11v2>(max_optin, true);
    tile_optin_shift<tile::Cfg<6, 2>>(max_optin, false);
    tile_optin_shift<tile::Cfg<7, 2>>(max_optin, false);
    tile_optin_shift<tile::Cfg<8, 2>>(max_optin, false);
    tile_optin_shift<tile::Cfg<9, 2>>(max_optin, false);
    tile_optin_shift<tile::Cfg<10, 2>>(max_optin, false);
    tile_optin_shift<tile::Cfg<11, 2>>(max_optin, false);
}

// tiles per CTA: long enough to amortise the twiddle prologue, short enough to keep >= ~8 waves of CTAs in flight
static size_t pick_tiles_per_cta(const ola_ctx* ctx, size_t total_tiles, size_t per_table, int resident_per_sm) {
    const size_t slots = (size_t)ctx->sm_count * (size_t)resident_per_sm;
    size_t t = total_tiles / (8 * slots);
    t = std::min<size_t>(std::max<size_t>(t, 1), 32);
    return std::min(t, per_table);
}
template <typename G, bool GS, int MODE = 0>
static void tile_strided_launch(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    PassArgs b = a;
    b.ncols = ncols;
    const size_t nsub = (size_t)1 << (a.L - a.M);
    const size_t per_q = ncols * ((((size_t)1 << a.M) >> G::l) / G::C);
    b.tiles_per_cta = pick_tiles_per_cta(ctx, per_q * nsub * (size_t)ncosets, per_q, G::MINB);
    size_t chunks = (per_q + b.tiles_per_cta - 1) / b.tiles_per_cta;
    while (nsub * chunks > 65535) {  // grid.y limit
        b.tiles_per_cta *= 2;
        chunks = (per_q + b.tiles_per_cta - 1) / b.tiles_per_cta;
    }
    const dim3 g((unsigned)ncosets, (unsigned)(nsub * chunks));
    tile::tile_strided<G, GS, MODE><<<g, G::NT, G::SMEM, ctx->stream>>>(b);
}
template <typename G, bool GS, int MODE = 0, bool FULL = false>
static void tile_contig_launch(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    PassArgs b = a;
    b.ncols = ncols;
    b.prefetch = tune_prefetch();
    const size_t nsub = ((size_t)1 << a.L) >> G::l;
    const size_t groups = (ncols + G::C - 1) / G::C;
    b.tiles_per_cta = pick_tiles_per_cta(ctx, groups * nsub * (size_t)ncosets, groups, G::MINB);
    size_t chunks = (groups + b.tiles_per_cta - 1) / b.tiles_per_cta;
    while (chunks > 65535) {
        b.tiles_per_cta *= 2;
        chunks = (groups + b.tiles_per_cta - 1) / b.tiles_per_cta;
    }
    const dim3 g((unsigned)nsub, (unsigned)chunks, (unsigned)ncosets);
    tile::tile_contig<G, GS, MODE, FULL><<<g, G::NT, G::SMEM, ctx->stream>>>(b);
}
// the default configurations also exist in the shift form (forward network; MODE 1 / 2 = forward / inverse roots)
template <typename G, bool GS>
static void tile_strided_launch_m(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    if constexpr (!GS) {
        if (tune_shift()) {
            if (a.inv_roots) tile_strided_launch<G, false, 2>(ctx, a, ncols, ncosets);
            else tile_strided_launch<G, false, 1>(ctx, a, ncols, ncosets);
            return;
        }
    }
    tile_strided_launch<G, GS>(ctx, a, ncols, ncosets);
}
template <typename G, bool GS>
static void tile_contig_launch_m(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    if constexpr (!GS) {
        if (tune_shift()) {
            const bool full = (ncols % G::C) == 0;  // no partial column group: the variant without per-lane predicates
            if (a.inv_roots) {
                if (full) tile_contig_launch<G, false, 2, true>(ctx, a, ncols, ncosets);
                else tile_contig_launch<G, false, 2>(ctx, a, ncols, ncosets);
            } else {
                if (full) tile_contig_launch<G, false, 1, true>(ctx, a, ncols, ncosets);
                else tile_contig_launch<G, false, 1>(ctx, a, ncols, ncosets);
            }
            return;
        }
    }
    tile_contig_launch<G, GS>(ctx, a, ncols, ncosets);
}
template <bool GS>
static void tile_strided_dispatch(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    const int v = tune_variant();
    switch (a.l) {
        case 6: tile_strided_launch_m<T6, GS>(ctx, a, ncols, ncosets); break;
        case 7: tile_strided_launch_m<T7, GS>(ctx, a, ncols, ncosets); break;
        case 8: tile_strided_launch_m<T8, GS>(ctx, a, ncols, ncosets); break;
        case 9: tile_strided_launch_m<T9, GS>(ctx, a, ncols, ncosets); break;
        case 10:
            if (v == 1) tile_strided_launch<T10v1, GS>(ctx, a, ncols, ncosets);
            else if (v == 2 && tune_strided_c4()) tile_strided_launch_m<T10c4, GS>(ctx, a, ncols, ncosets);
            else if (v == 2) tile_strided_launch_m<T10v2, GS>(ctx, a, ncols, ncosets);
            else if (v == 3) tile_strided_launch<T10v3, GS>(ctx, a, ncols, ncosets);
            else tile_strided_launch<T10, GS>(ctx, a, ncols, ncosets);
            break;
        case 11:
            if (v == 2) tile_strided_launch_m<T11v2, GS>(ctx, a, ncols, ncosets);
            else if (v == 3) tile_strided_launch<T11v3, GS>(ctx, a, ncols, ncosets);
            else tile_strided_launch<T11, GS>(ctx, a, ncols, ncosets);
            break;
        default: OLA_CHECK(false, OLA_ERR_INTERNAL, "no tiled strided pass of that length");
    }
}
template <bool GS>
static void tile_contig_dispatch(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    const int v = tune_variant();
    if (ncols <= 2) {  // one or two columns (quotient, FRI): 2-lane tiles
        switch (a.l) {
            case 6: tile_contig_launch_m<tile::Cfg<6, 2>, GS>(ctx, a, ncols, ncosets); break;
            case 7: tile_contig_launch_m<tile::Cfg<7, 2>, GS>(ctx, a, ncols, ncosets); break;
            case 8: tile_contig_launch_m<tile::Cfg<8, 2>, GS>(ctx, a, ncols, ncosets); break;
            case 9: tile_contig_launch_m<tile::Cfg<9, 2>, GS>(ctx, a, ncols, ncosets); break;
            case 10: tile_contig_launch_m<tile::Cfg<10, 2>, GS>(ctx, a, ncols, ncosets); break;
            case 11: tile_contig_launch_m<tile::Cfg<11, 2>, GS>(ctx, a, ncols, ncosets); break;
            default: OLA_CHECK(false, OLA_ERR_INTERNAL, "no tiled contiguous pass of that length");
        }
        return;
    }
    switch (a.l) {
        case 6: tile_contig_launch_m<T6, GS>(ctx, a, ncols, ncosets); break;
        case 7: tile_contig_launch_m<T7, GS>(ctx, a, ncols, ncosets); break;
        case 8: tile_contig_launch_m<T8, GS>(ctx, a, ncols, ncosets); break;
        case 9: tile_contig_launch_m<T9, GS>(ctx, a, ncols, ncosets); break;
        case 10:
            if (tune_contig_c4()) tile_contig_launch_m<T10c4, GS>(ctx, a, ncols, ncosets);
            else if (v == 1) tile_contig_launch<T10v1, GS>(ctx, a, ncols, ncosets);
            else if (v == 2) tile_contig_launch<T10v2, GS>(ctx, a, ncols, ncosets);
            else if (v == 3) tile_contig_launch<T10v3, GS>(ctx, a, ncols, ncosets);
            else tile_contig_launch<T10, GS>(ctx, a, ncols, ncosets);
            break;
        case 11:
            if (v == 2) tile_contig_launch_m<T11v2, GS>(ctx, a, ncols, ncosets);
            else if (v == 3) tile_contig_launch<T11v3, GS>(ctx, a, ncols, ncosets);
            else tile_contig_launch<T11, GS>(ctx, a, ncols, ncosets);
            break;
        default: OLA_CHECK(false, OLA_ERR_INTERNAL, "no tiled contiguous pass of that length");
    }
}

// natural-order last pass (tile_nat): C sub-blocks per tile; needs at least C sub-blocks (t = L - l >= log2 C)
using N6 = tile::Cfg<6, 8, 2, 1, 4, 1>;
using N7 = tile::Cfg<7, 8, 2, 1, 4, 1>;
using N8 = tile::Cfg<8, 8, 2, 1, 4, 1>;
using N9 = tile::Cfg<9, 8, 2, 1, 4, 1>;
using N10 = tile::Cfg<10, 8, 2, 1, 4, 1>;
using N11 = tile::Cfg<11, 4, 2, 1, 4, 1>;
// four sub-blocks per tile (whole 32-byte sectors per store), 256 threads, register cap and 68 KB of shared memory for 3 CTAs / SM:
// 1.43 ms against 1.95 ms for the eight-lane tile at one 512-thread CTA / SM (profiles/r03h_nat_c4_ab.txt); OLA_NTT_NAT_C4=0 selects the latter
using N10c4 = tile::Cfg<10, 4, 2, 1, 4, 3>;
static bool tune_nat_c4() {
    static bool v = [] {
        const char* e = getenv("OLA_NTT_NAT_C4");
        return !(e && atoi(e) == 0);
    }();
    return v;
}
template <typename G>
static constexpr size_t tile_nat_smem() {
    return ((size_t)G::C * G::R + (size_t)G::C * 16 + 16 + (size_t)G::RP * G::C) * sizeof(uint64_t);
}
template <typename G>
static void tile_nat_launch(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets) {
    PassArgs b = a;
    b.ncols = ncols;
    b.prefetch = tune_prefetch();
    const size_t groups = (((size_t)1 << a.L) >> G::l) / G::C;
    b.tiles_per_cta = pick_tiles_per_cta(ctx, groups * ncols * (size_t)ncosets, ncols, G::MINB);
    size_t chunks = (ncols + b.tiles_per_cta - 1) / b.tiles_per_cta;
    while (chunks > 65535) {
        b.tiles_per_cta *= 2;
        chunks = (ncols + b.tiles_per_cta - 1) / b.tiles_per_cta;
    }
    const dim3 g((unsigned)groups, (unsigned)chunks, (unsigned)ncosets);
    if (a.inv_roots) tile::tile_nat<G, 2><<<g, G::NT, tile_nat_smem<G>(), ctx->stream>>>(b);
    else tile::tile_nat<G, 1><<<g, G::NT, tile_nat_smem<G>(), ctx->stream>>>(b);
}
static bool tile_nat_dispatch(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets, const char* name) {
    const int t = a.L - a.l;
    if (t < 3) return false;  // fewer than eight sub-blocks: the generic pass handles it
    Launch lz(ctx, name);
    switch (a.l) {
        case 6: tile_nat_launch<N6>(ctx, a, ncols, ncosets); break;
        case 7: tile_nat_launch<N7>(ctx, a, ncols, ncosets); break;
        case 8: tile_nat_launch<N8>(ctx, a, ncols, ncosets); break;
        case 9: tile_nat_launch<N9>(ctx, a, ncols, ncosets); break;
        case 10:
            if (tune_nat_c4()) tile_nat_launch<N10c4>(ctx, a, ncols, ncosets);
            else tile_nat_launch<N10>(ctx, a, ncols, ncosets);
            break;
        default: tile_nat_launch<N11>(ctx, a, ncols, ncosets); break;
    }
    return true;
}
template <typename G>
static void tile_nat_optin(int max_optin) {
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_nat<G, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(tile::tile_nat<G, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
}

// ---------------------------------------------------------------------------------------------------
// Opt every pass kernel into the device's maximum dynamic shared memory once, with the same value from every context
// (a per-launch cudaFuncSetAttribute with the launch's own size races when several host threads drive one GPU).
static void opt_in_shared_memory(int device) {
    int max_optin = 0;
    OLA_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    OLA_CUDA(cudaFuncSetAttribute(pass_strided_r8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_strided_r8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_contig_r8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_contig_r8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_strided<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_strided<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_contig<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    OLA_CUDA(cudaFuncSetAttribute(pass_contig<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    tile_optin_all(max_optin);
    tile_nat_optin<N6>(max_optin);
    tile_nat_optin<N7>(max_optin);
    tile_nat_optin<N8>(max_optin);
    tile_nat_optin<N9>(max_optin);
    tile_nat_optin<N10>(max_optin);
    tile_nat_optin<N11>(max_optin);
    tile_nat_optin<N10c4>(max_optin);
}

void init_twiddles(ola_ctx* ctx) {
    opt_in_shared_memory(ctx->device);
    for (int d = 0; d < 2; ++d) {
        OLA_CUDA(cudaMalloc(&ctx->tw.pw[d], 5120 * sizeof(uint64_t)));
        OLA_CUDA(cudaMalloc(&ctx->tw.brs[d], 2048 * sizeof(uint64_t)));
        uint64_t omega = gl::TWO_ADIC_GEN, w12 = gl::root_of_unity(12);
        if (d == 1) {
            omega = gl::inv(omega);
            w12 = gl::inv(w12);
        }
        build_tables_kernel<<<8, 256, 0, ctx->stream>>>(ctx->tw.pw[d], ctx->tw.brs[d], omega, w12);
        check_launch("build_tables_kernel");
        count_launch(ctx);
    }
}
void free_twiddles(ola_ctx* ctx) {
    for (int d = 0; d < 2; ++d) {
        if (ctx->tw.pw[d]) cudaFree(ctx->tw.pw[d]);
        if (ctx->tw.brs[d]) cudaFree(ctx->tw.brs[d]);
        ctx->tw.pw[d] = ctx->tw.brs[d] = nullptr;
    }
}

std::vector<int> plan_passes(int L) {
    std::vector<int> p;
    if (L <= MAX_PASS_LOG) {
        p.push_back(L);
        return p;
    }
    int k = (L + MAX_PASS_LOG - 1) / MAX_PASS_LOG;
    int base = L / k, rem = L % k;
    for (int i = 0; i < k; ++i) p.push_back(base + (i < rem ? 1 : 0));
    return p;
}

static uint64_t pow2k(uint64_t x, int k) {  // x^(2^k)
    for (int i = 0; i < k; ++i) x = gl::sqr(x);
    return x;
}

// tuning knobs (environment overrides are for profiling sessions only)
static int tune_threads() {
    static int v = [] {
        const char* e = getenv("OLA_NTT_THREADS");
        int t = e ? atoi(e) : 512;
        return (t >= 32 && t <= 1024) ? t : 512;
    }();
    return v;
}
static bool tune_tile() {  // OLA_NTT_TILE=0 falls back to the generic passes (A/B measurements)
    static bool v = [] {
        const char* e = getenv("OLA_NTT_TILE");
        return !(e && atoi(e) == 0);
    }();
    return v;
}
static int tune_gmax() {
    static int v = [] {
        const char* e = getenv("OLA_NTT_G");
        int t = e ? atoi(e) : 2;  // profiles/r01d_ntt_sweep.txt: G = 2, 512 threads is the fastest pair
        return (t == 1 || t == 2 || t == 4 || t == 8) ? t : 2;
    }();
    return v;
}

template <bool GS>
static void launch_strided(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets, const char* name) {
    const int R = 1 << a.l;
    size_t tiles = ((size_t)1 << a.L) / ((size_t)R * TILE_T);
    dim3 grid((unsigned)tiles, (unsigned)ncols, (unsigned)ncosets);
    if (a.l >= 6 && a.l <= 11 && tune_tile()) {
        Launch lz(ctx, name);
        tile_strided_dispatch<GS>(ctx, a, ncols, ncosets);
    } else if (a.l >= 6) {
        const size_t padded = (size_t)R * TILE_T + ((size_t)R * TILE_T >> 4) + 1;
        size_t smem = ((size_t)R + 16 + padded) * sizeof(uint64_t);
        int threads = (int)std::min<size_t>((size_t)tune_threads(), (size_t)R);  // R items of 8 elements per round
        Launch lz(ctx, name);
        pass_strided_r8<GS><<<grid, threads, smem, ctx->stream>>>(a);
    } else {
        size_t smem = ((size_t)R + 16 + (size_t)R * TILE_T) * sizeof(uint64_t);
        int threads = (int)std::min<size_t>(1024, std::max<size_t>(64, (size_t)R * TILE_T / 4));
        Launch lz(ctx, name);
        pass_strided<GS><<<grid, threads, smem, ctx->stream>>>(a);
    }
    check_launch("pass_strided");
}

template <bool GS>
static void launch_contig(ola_ctx* ctx, const PassArgs& a, size_t ncols, int ncosets, const char* name) {
    const int R = 1 << a.l;
    size_t blocks = (((size_t)1 << a.L) >> a.l) / a.G;
    dim3 grid((unsigned)blocks, (unsigned)ncols, (unsigned)ncosets);
    if (a.bitrev_store && a.l >= 6 && a.l <= 11 && tune_tile() && tune_shift() && tune_nat() == 2 && tile_nat_dispatch(ctx, a, ncols, ncosets, name)) {
        // done by tile_nat
    } else if (a.l >= 6 && a.l <= 11 && !a.bitrev_store && tune_tile()) {
        Launch lz(ctx, name);
        tile_contig_dispatch<GS>(ctx, a, ncols, ncosets);
    } else if (a.l >= 6) {
        const size_t padded = (size_t)a.G * R + ((size_t)a.G * R >> 4) + 1;
        size_t smem = ((size_t)a.G * R + (size_t)a.G * 16 + padded) * sizeof(uint64_t);
        int threads = (int)std::min<size_t>((size_t)tune_threads(), std::max<size_t>(32, (size_t)R * a.G / 8));
        Launch lz(ctx, name);
        pass_contig_r8<GS><<<grid, threads, smem, ctx->stream>>>(a);
    } else {
        size_t smem = ((size_t)2 * a.G * R + (size_t)a.G * 16) * sizeof(uint64_t);
        int threads = (int)std::min<size_t>(1024, std::max<size_t>(32, (size_t)R * a.G / 4));
        Launch lz(ctx, name);
        pass_contig<GS><<<grid, threads, smem, ctx->stream>>>(a);
    }
    check_launch("pass_contig");
}

// Forward network (natural-order input -> bit-reversed positions, or natural order if bitrev_store).
void forward(ola_ctx* ctx, const FwdDesc& d) {
    OLA_CHECK(d.log_n <= 32 - d.coset_bits, OLA_ERR_INVALID_ARG, "NTT size exceeds the field's two-adicity (2^32)");
    OLA_CHECK((1 << d.coset_bits) <= MAX_COSETS, OLA_ERR_INVALID_ARG, "too many cosets");
    if (d.ncols == 0) return;
    const int L = d.log_n, all_cosets = 1 << d.coset_bits;
    const int ncosets = d.coset_count < 0 ? all_cosets : d.coset_count;
    OLA_CHECK(d.coset_first >= 0 && ncosets >= 1 && d.coset_first + ncosets <= all_cosets, OLA_ERR_INVALID_ARG, "coset range out of bounds");
    OLA_CHECK(!(d.natural_output && ncosets != all_cosets), OLA_ERR_INVALID_ARG, "natural-order LDE output needs every coset");
    const int dir = d.inverse_roots ? 1 : 0;
    // coset i evaluates on shift * g^{bitrev(i)} * H_n, g = omega_{n * ncosets}  (cfft/serial.rs:36-38)
    uint64_t g = gl::root_of_unity(L + d.coset_bits);
    if (d.inverse_roots) g = gl::inv(g);
    uint64_t shifts[MAX_COSETS];
    for (int i = 0; i < ncosets; ++i) shifts[i] = gl::mul(d.shift, gl::pow(g, gl::bitrev32((uint32_t)(d.coset_first + i), d.coset_bits)));

    std::vector<int> plan = plan_passes(L);
    int M = L;
    for (size_t pi = 0; pi < plan.size(); ++pi) {
        PassArgs a{};
        const bool first = (pi == 0), last = (pi + 1 == plan.size());
        uint64_t* work = d.work ? d.work : d.dst;
        OLA_CHECK(!(d.natural_output && plan.size() > 1 && work == d.dst), OLA_ERR_INTERNAL,
                  "natural-order output of a multi-pass transform needs a work buffer distinct from dst");
        // first pass reads src; intermediate passes run in `work`; the last pass writes dst
        a.src = first ? d.src : work;
        a.dst = last ? d.dst : work;
        const size_t work_col = d.work ? d.work_col_stride : d.dst_col_stride;
        const size_t work_coset = d.work ? d.work_coset_stride : d.dst_coset_stride;
        a.src_col_stride = first ? d.src_col_stride : work_col;
        a.dst_col_stride = last ? d.dst_col_stride : work_col;
        a.src_coset_stride = first ? 0 : work_coset;
        a.dst_coset_stride = last ? d.dst_coset_stride : work_coset;
        a.pw = ctx->tw.pw[dir];
        a.brs = ctx->tw.brs[dir];
        a.inv_roots = dir;
        a.L = L;
        a.M = M;
        a.l = plan[pi];
        a.apply_scale = (last && d.apply_scale) ? 1 : 0;
        a.lazy_out = last ? 0 : 1;
        a.scale = d.scale;
        a.coset_bits = 0;
        a.out_mul = 1;
        for (int i = 0; i < ncosets; ++i) a.s_last[i] = pow2k(shifts[i], M - a.l);
        if (!last) {
            launch_strided<false>(ctx, a, d.ncols, ncosets, d.tag_strided);
        } else {
            const int t = L - M;
            int gmax = (a.l <= 10) ? tune_gmax() : std::min(tune_gmax(), 4);
            a.G = (int)std::min<size_t>((size_t)gmax, (size_t)1 << t);
            a.bitrev_store = d.natural_output ? 1 : 0;
            if (d.natural_output) {
                a.coset_bits = d.coset_bits;
                a.out_mul = (size_t)ncosets;
            }
            launch_contig<false>(ctx, a, d.ncols, ncosets, d.tag_contig);
        }
        M -= a.l;
    }
}

// Inverse network: bit-reversed positions on the coset shift*H (leaf order) -> natural coefficients, x 1/n.
void inverse_from_leaf_order(ola_ctx* ctx, uint64_t* data, size_t col_stride, size_t ncols, int log_n, uint64_t shift) {
    OLA_CHECK(log_n <= 32, OLA_ERR_INVALID_ARG, "NTT size exceeds the field's two-adicity (2^32)");
    if (ncols == 0) return;
    const int L = log_n;
    std::vector<int> plan = plan_passes(L);
    const uint64_t shift_inv = gl::inv(shift);
    const uint64_t n_inv = gl::inv(((uint64_t)1 << L) % gl::P);
    // undo passes in reverse order: the last forward pass (contiguous) first
    std::vector<int> Ms(plan.size());
    int M = L;
    for (size_t pi = 0; pi < plan.size(); ++pi) {
        Ms[pi] = M;
        M -= plan[pi];
    }
    for (size_t k = plan.size(); k-- > 0;) {
        PassArgs a{};
        a.src = data;
        a.dst = data;
        a.src_col_stride = a.dst_col_stride = col_stride;
        a.pw = ctx->tw.pw[1];
        a.brs = ctx->tw.brs[1];
        a.L = L;
        a.M = Ms[k];
        a.l = plan[k];
        a.apply_scale = (k == 0) ? 1 : 0;
        a.lazy_out = (k == 0) ? 0 : 1;
        a.scale = n_inv;
        a.out_mul = 1;
        a.s_last[0] = pow2k(shift_inv, a.M - a.l);
        if (k + 1 == plan.size()) {
            const int t = L - a.M;
            int gmax = (a.l <= 10) ? tune_gmax() : std::min(tune_gmax(), 4);
            a.G = (int)std::min<size_t>((size_t)gmax, (size_t)1 << t);
            launch_contig<true>(ctx, a, ncols, 1, "coset_intt_contig");
        } else {
            launch_strided<true>(ctx, a, ncols, 1, "coset_intt_strided");
        }
    }
}

void scale_powers(ola_ctx* ctx, uint64_t* data, size_t col_stride, size_t ncols, size_t n, uint64_t base, uint64_t step) {
    if (ncols == 0 || n == 0) return;
    size_t runs = (n + 15) / 16;
    dim3 grid((unsigned)((runs + 127) / 128), (unsigned)ncols);
    scale_powers_kernel<<<grid, 128, 0, ctx->stream>>>(data, col_stride, n, base, step);
    check_launch("scale_powers_kernel");
    count_launch(ctx);
}

}  // namespace ntt
}  // namespace ola
