// Register-tiled NTT passes (included by ntt.cu): the hot kernels of the forward (coset-LDE) and inverse networks.
//
// Same butterfly network, twiddles and data layout as the generic passes in ntt.cu (so the two families can be mixed
// pass by pass), rebuilt around the instruction budget: a Goldilocks butterfly is ~28 SASS instructions of arithmetic
// (gl.cuh) and everything else in the kernel is overhead to be amortised.
//   * A tile is R = 2^l transform rows x C "lanes" that SHARE twiddles: C adjacent inner indices in a strided pass
//     (one 64-byte segment per row), C columns of the batch in a contiguous pass.  The per-tile twiddle table
//     (R - 1 products) is therefore built once per R*C elements, and every twiddle fetched from shared memory feeds
//     two butterflies (each thread owns the same rows of two adjacent lanes).
//   * The l stages of a pass run as radix-8 / radix-16 register rounds (3 or 4 stages on 8 or 16 rows x 2 lanes held in
//     registers): ceil(l / 4) rounds, shared memory is touched once between rounds with 128-bit accesses, the first
//     round loads straight from HBM and the last one stores straight back.  The radix-16 rounds come last so the
//     stride-1 round is radix-16 whenever one exists.
//   * Everything about a round (strides, twiddle offsets, loop trip counts) is a compile-time constant of <l, round>;
//     per item the only address arithmetic is one base pointer per array.
//   * Twiddles are stored per round as [twiddle-of-the-round][qh] so that the stride-1 round (qh = thread) reads them
//     conflict-free and the other rounds broadcast.  Rows are padded by one row per 16 so every round's 128-bit accesses
//     are bank-conflict-free.
//   * Inputs may be any u64 representatives ("lazy"); only the last pass of a transform canonicalises its output.
#pragma once

namespace ola {
namespace ntt {
namespace tile {

template <int l>
struct Sched {
    static constexpr int NR = (l + 3) / 4;    // rounds
    static constexpr int N16 = l - 3 * NR;    // radix-16 rounds (the last N16)
    static constexpr int N8 = NR - N16;       // radix-8 rounds (the first N8)
    static_assert(N16 >= 0 && N8 >= 0, "unsupported pass length");
};

template <int l, int RHO>
struct Rd {
    static constexpr int K = RHO < Sched<l>::N8 ? 3 : 4;  // stages of this round
    static constexpr int U0 = RHO < Sched<l>::N8 ? 3 * RHO : 3 * Sched<l>::N8 + 4 * (RHO - Sched<l>::N8);  // first stage
    static constexpr int SH = l - U0 - K;                  // log2 of the row stride inside a register block
    static constexpr int OFF = Rd<l, RHO - 1>::OFF + (((1 << Rd<l, RHO - 1>::K) - 1) << Rd<l, RHO - 1>::U0);
};
template <int l>
struct Rd<l, 0> {
    static constexpr int K = 0 < Sched<l>::N8 ? 3 : 4;
    static constexpr int U0 = 0;
    static constexpr int SH = l - K;
    static constexpr int OFF = 0;
};

template <int l, int C>
struct Geo {
    static constexpr int R = 1 << l;
    static constexpr int RP = R + (R >> 4);  // padded rows
    static constexpr int KMAX = Sched<l>::N16 > 0 ? 4 : 3;
    static constexpr int NT_RAW = (R >> KMAX) * (C / 2);
    static constexpr int NT = NT_RAW < 32 ? 32 : (NT_RAW > 512 ? 512 : NT_RAW);
    static constexpr size_t SMEM = ((size_t)R + 16 + (size_t)RP * C) * sizeof(uint64_t);
};

// twiddle of (stage u0 + s, block (qh << s) + ql) lives at tw[OFF + (((1 << s) - 1 + ql) << U0) + qh]
template <int l, int RHO>
__device__ __forceinline__ void scatter_twiddle(uint64_t* tw, int u, int q, uint64_t v) {
    using rd = Rd<l, RHO>;
    if (u >= rd::U0 && u < rd::U0 + rd::K) {
        const int s = u - rd::U0;
        const int qh = q >> s, ql = q & ((1 << s) - 1);
        tw[rd::OFF + ((((1 << s) - 1) + ql) << rd::U0) + qh] = v;
    }
    if constexpr (RHO + 1 < Sched<l>::NR) scatter_twiddle<l, RHO + 1>(tw, u, q, v);
}

template <int l>
__device__ __forceinline__ void build_twiddles(uint64_t* tw, const uint64_t* cu, const uint64_t* __restrict__ brs, int tid, int nt) {
    for (int i = tid + 1; i < (1 << l); i += nt) {
        const int u = 31 - __clz(i), q = i - (1 << u);
        scatter_twiddle<l, 0>(tw, u, q, gl::mul(cu[u], __ldg(brs + q)));
    }
}

// K stages on 2^K rows x 2 lanes in registers; t = tw + OFF + qh
template <int K, int U0, bool GS>
__device__ __forceinline__ void bfly_regs(uint64_t (&v)[1 << K][2], const uint64_t* __restrict__ t) {
    if (!GS) {
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
                    for (int ln = 0; ln < 2; ++ln) {
                        const uint64_t p = gl::canon_fast(gl::mul_lazy(v[m + half][ln], w));
                        const uint64_t a = v[m][ln];
                        v[m][ln] = gl::add_lc(a, p);
                        v[m + half][ln] = gl::sub_lc(a, p);
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int s = K - 1; s >= 0; --s) {
            const int half = (1 << (K - 1)) >> s;
#pragma unroll
            for (int ql = 0; ql < (1 << s); ++ql) {
                const uint64_t w = t[(((1 << s) - 1) + ql) << U0];
#pragma unroll
                for (int jj = 0; jj < half; ++jj) {
                    const int m = (ql << (K - s)) + jj;
#pragma unroll
                    for (int ln = 0; ln < 2; ++ln) {
                        const uint64_t b = gl::canon_fast(v[m + half][ln]);
                        const uint64_t a = v[m][ln];
                        v[m][ln] = gl::add_lc(a, b);
                        v[m + half][ln] = gl::mul_lazy(gl::sub_lc(a, b), w);
                    }
                }
            }
        }
    }
}

// global-memory view of a tile
template <bool CONTIG>
struct Io {
    const uint64_t* in;   // element (row 0, lane 0)
    uint64_t* out;
    size_t in_row, out_row;    // strided: elements between consecutive rows (2^(M-l)); contig: 1
    size_t in_lane, out_lane;  // strided: 1; contig: column stride
    int lanes_valid;           // contig: columns of this group that exist (lanes >= lanes_valid are skipped)
    int apply_scale, lazy_out;
    uint64_t scale;
};

template <int SH>
__device__ __forceinline__ constexpr int pad_of(int m) {
    return SH >= 4 ? (m << (SH >= 4 ? SH - 4 : 0)) : (m >> (SH < 4 ? 4 - SH : 0));
}

template <int l, int C, bool GS, bool CONTIG, int I>
__device__ __forceinline__ void run_step(const Io<CONTIG>& io, uint64_t* __restrict__ x, const uint64_t* __restrict__ tw, int tid) {
    constexpr int NR = Sched<l>::NR;
    constexpr int RHO = GS ? NR - 1 - I : I;
    using rd = Rd<l, RHO>;
    constexpr int K = rd::K, U0 = rd::U0, SH = rd::SH, NE = 1 << K;
    constexpr bool FIRST = (I == 0), LAST = (I == NR - 1);
    constexpr int NITEMS = ((1 << l) >> K) * (C / 2);
    constexpr int NT = Geo<l, C>::NT;
#pragma unroll 1
    for (int w = tid; w < NITEMS; w += NT) {
        const int cp = w % (C / 2), rest = w / (C / 2);
        const int j = rest & ((1 << SH) - 1), qh = rest >> SH;
        const int rbase = (qh << (SH + K)) + j;
        uint64_t v[NE][2];
        ulonglong2* xp = reinterpret_cast<ulonglong2*>(x + (size_t)(rbase + (rbase >> 4)) * C + 2 * cp);
        if (FIRST) {
            if (!CONTIG) {
                const uint64_t* p = io.in + (size_t)rbase * io.in_row + 2 * cp;
                const size_t step = io.in_row << SH;
#pragma unroll
                for (int m = 0; m < NE; ++m) {
                    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p + (size_t)m * step);
                    v[m][0] = q.x;
                    v[m][1] = q.y;
                }
            } else {
#pragma unroll
                for (int ln = 0; ln < 2; ++ln) {
                    const bool ok = 2 * cp + ln < io.lanes_valid;
                    const uint64_t* p = io.in + (size_t)(2 * cp + ln) * io.in_lane + rbase;
                    if (SH == 0) {
#pragma unroll
                        for (int m = 0; m < NE; m += 2) {
                            ulonglong2 q = make_ulonglong2(0, 0);
                            if (ok) q = *reinterpret_cast<const ulonglong2*>(p + m);
                            v[m][ln] = q.x;
                            v[m + 1][ln] = q.y;
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < NE; ++m) v[m][ln] = ok ? p[(size_t)m << SH] : 0;
                    }
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < NE; ++m) {
                const ulonglong2 q = xp[(((m << SH) + pad_of<SH>(m)) * C) / 2];
                v[m][0] = q.x;
                v[m][1] = q.y;
            }
        }
        bfly_regs<K, U0, GS>(v, tw + rd::OFF + qh);
        if (LAST) {
#pragma unroll
            for (int m = 0; m < NE; ++m) {
#pragma unroll
                for (int ln = 0; ln < 2; ++ln) {
                    if (io.apply_scale)
                        v[m][ln] = gl::mul(v[m][ln], io.scale);
                    else if (!io.lazy_out)
                        v[m][ln] = gl::canon_fast(v[m][ln]);
                }
            }
            if (!CONTIG) {
                uint64_t* p = io.out + (size_t)rbase * io.out_row + 2 * cp;
                const size_t step = io.out_row << SH;
#pragma unroll
                for (int m = 0; m < NE; ++m) *reinterpret_cast<ulonglong2*>(p + (size_t)m * step) = make_ulonglong2(v[m][0], v[m][1]);
            } else {
#pragma unroll
                for (int ln = 0; ln < 2; ++ln) {
                    if (2 * cp + ln >= io.lanes_valid) continue;
                    uint64_t* p = io.out + (size_t)(2 * cp + ln) * io.out_lane + rbase;
                    if (SH == 0) {
#pragma unroll
                        for (int m = 0; m < NE; m += 2) *reinterpret_cast<ulonglong2*>(p + m) = make_ulonglong2(v[m][ln], v[m + 1][ln]);
                    } else {
#pragma unroll
                        for (int m = 0; m < NE; ++m) p[(size_t)m << SH] = v[m][ln];
                    }
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < NE; ++m) xp[(((m << SH) + pad_of<SH>(m)) * C) / 2] = make_ulonglong2(v[m][0], v[m][1]);
        }
    }
    if (!LAST) __syncthreads();
}

template <int l, int C, bool GS, bool CONTIG, int I>
__device__ __forceinline__ void run_steps(const Io<CONTIG>& io, uint64_t* x, const uint64_t* tw, int tid) {
    run_step<l, C, GS, CONTIG, I>(io, x, tw, tid);
    if constexpr (I + 1 < Sched<l>::NR) run_steps<l, C, GS, CONTIG, I + 1>(io, x, tw, tid);
}

// strided pass: tile = (sub-block Q, C adjacent inner indices) x all 2^l rows; grid (tiles, columns, cosets)
template <int l, int C, bool GS>
__global__ void __launch_bounds__(Geo<l, C>::NT) tile_strided(const PassArgs a) {
    extern __shared__ __align__(16) uint64_t sm[];
    constexpr int R = 1 << l;
    uint64_t* tw = sm;
    uint64_t* cu = sm + R;
    uint64_t* x = sm + R + 16;
    const int tid = threadIdx.x;
    // cosets vary fastest in the grid: the cosets of one tile are resident together, so the shared input of a coset LDE
    // (src_coset_stride == 0) comes from HBM once and from L2 for the other cosets
    const uint32_t coset = a.coset_major ? blockIdx.x : blockIdx.z;
    const uint32_t tile_id = a.coset_major ? blockIdx.y : blockIdx.x;
    const uint32_t col = a.coset_major ? blockIdx.z : blockIdx.y;
    const size_t inner = (size_t)1 << (a.M - l);
    const uint32_t tiles_per_sub = (uint32_t)(inner / C);
    const uint32_t Q = tile_id / tiles_per_sub;
    const size_t c0 = (size_t)(tile_id % tiles_per_sub) * C;
    Io<false> io;
    io.in = a.src + col * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << a.M) + c0;
    io.out = a.dst + col * a.dst_col_stride + coset * a.dst_coset_stride + ((size_t)Q << a.M) + c0;
    io.in_row = io.out_row = inner;
    io.in_lane = io.out_lane = 1;
    io.lanes_valid = C;
    io.apply_scale = a.apply_scale;
    io.lazy_out = a.lazy_out;
    io.scale = a.scale;
    if (tid == 0) row_constants(a, a.s_last[coset], Q, cu);
    __syncthreads();
    build_twiddles<l>(tw, cu, a.brs, tid, Geo<l, C>::NT);
    __syncthreads();
    run_steps<l, C, GS, false, 0>(io, x, tw, tid);
}

// contiguous pass (M == l): tile = sub-block Q (2^l consecutive elements) of C columns; grid (sub-blocks, column groups, cosets)
template <int l, int C, bool GS>
__global__ void __launch_bounds__(Geo<l, C>::NT) tile_contig(const PassArgs a) {
    extern __shared__ __align__(16) uint64_t sm[];
    constexpr int R = 1 << l;
    uint64_t* tw = sm;
    uint64_t* cu = sm + R;
    uint64_t* x = sm + R + 16;
    const int tid = threadIdx.x;
    const uint32_t coset = blockIdx.z;
    const uint32_t Q = blockIdx.x;
    const size_t col0 = (size_t)blockIdx.y * C;
    Io<true> io;
    io.in = a.src + col0 * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << l);
    io.out = a.dst + col0 * a.dst_col_stride + coset * a.dst_coset_stride + ((size_t)Q << l);
    io.in_row = io.out_row = 1;
    io.in_lane = a.src_col_stride;
    io.out_lane = a.dst_col_stride;
    io.lanes_valid = (int)((a.ncols - col0) < (size_t)C ? (a.ncols - col0) : (size_t)C);
    io.apply_scale = a.apply_scale;
    io.lazy_out = a.lazy_out;
    io.scale = a.scale;
    if (tid == 0) row_constants(a, a.s_last[coset], Q, cu);
    __syncthreads();
    build_twiddles<l>(tw, cu, a.brs, tid, Geo<l, C>::NT);
    __syncthreads();
    run_steps<l, C, GS, true, 0>(io, x, tw, tid);
}

}  // namespace tile
}  // namespace ntt
}  // namespace ola
