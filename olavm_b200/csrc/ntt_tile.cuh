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
#include "ntt_shift.cuh"
#include "tma.cuh"

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

// Tile shape and thread mapping.  LN3 / LN4 = adjacent lanes owned by one thread in radix-8 / radix-16 rounds (1 or 2);
// PADLOG: one padding row per 2^PADLOG rows; MINB: resident CTAs per SM the register allocation is capped for.
template <int l_, int C_, int LN3_ = 2, int LN4_ = 2, int PADLOG_ = 4, int MINB_ = 2>
struct Cfg {
    static constexpr int l = l_, C = C_, LN3 = LN3_, LN4 = LN4_, PADLOG = PADLOG_, MINB = MINB_;
    static constexpr int R = 1 << l;
    static constexpr int RP = R + (R >> PADLOG);  // padded rows
    static constexpr int ITEMS3 = (R >> 3) * (C / LN3), ITEMS4 = (R >> 4) * (C / LN4);
    static constexpr int NT_RAW = Sched<l>::N16 == 0 ? ITEMS3 : (Sched<l>::N8 == 0 ? ITEMS4 : (ITEMS3 < ITEMS4 ? ITEMS3 : ITEMS4));
    static constexpr int NT = NT_RAW < 32 ? 32 : (NT_RAW > 512 ? 512 : NT_RAW);
    static constexpr size_t SMEM = ((size_t)R + 16 + (size_t)RP * C) * sizeof(uint64_t);
    static_assert(C % LN3 == 0 && C % LN4 == 0 && (LN3 == 1 || LN3 == 2) && (LN4 == 1 || LN4 == 2), "lanes per thread");
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

// shift form (ntt_shift.cuh): the table of round RHO holds theta^m, m = 1 .. 2^K - 1, at tw[OFF + ((m - 1) << U0) + qh] with
// theta = the twiddle of (stage U0 + K - 1, block qh << (K - 1)) -- the same 2^K - 1 slots per qh as the radix-2 form
// (the LAST round's table is theta^m * scale when the pass applies an output scale: bfly_shift's scale0)
template <int l, int RHO>
__device__ __forceinline__ void build_twiddles_pow_round(uint64_t* tw, const uint64_t* cu, const uint64_t* __restrict__ brs, int tid, int nt, uint64_t scale) {
    using rd = Rd<l, RHO>;
    for (int qh = tid; qh < (1 << rd::U0); qh += nt) {
        const uint64_t theta = gl::mul(cu[rd::U0 + rd::K - 1], __ldg(brs + (qh << (rd::K - 1))));
        uint64_t p = (RHO + 1 == Sched<l>::NR) ? gl::mul(theta, scale) : theta;
        tw[rd::OFF + qh] = p;
#pragma unroll 1
        for (int m = 2; m < (1 << rd::K); ++m) {
            p = gl::mul(p, theta);
            tw[rd::OFF + ((m - 1) << rd::U0) + qh] = p;
        }
    }
    if constexpr (RHO + 1 < Sched<l>::NR) build_twiddles_pow_round<l, RHO + 1>(tw, cu, brs, tid, nt, scale);
}

// K stages on 2^K rows x LN lanes in registers; t = tw + OFF + qh
template <int K, int U0, bool GS, int LN>
__device__ __forceinline__ void bfly_regs(uint64_t (&v)[1 << K][LN], const uint64_t* __restrict__ t) {
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
                    for (int ln = 0; ln < LN; ++ln) {
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
                    for (int ln = 0; ln < LN; ++ln) {
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
    // contig, natural-order output: row r of the sub-block goes to out[lane * out_lane + ((bitrev_l(r) << br_t) * br_mul) + br_off]
    int br_t;                  // -1: positions stay as they are
    size_t br_mul, br_off;
    // tile_nat: the lanes are C sub-blocks (not columns): their input offsets come from a table and each has its own
    // twiddle table, twl elements after the previous lane's (0 = one table for all lanes)
    const size_t* lane_in;
    size_t twl;
};

__device__ __forceinline__ constexpr uint32_t brev_const(uint32_t m, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((m >> i) & 1u) << (bits - 1 - i);
    return r;
}

// row offset (in padded rows) of register-block element m relative to the block's first row
template <int SH, int P>
__device__ __forceinline__ constexpr int row_off(int m) {
    return (m << SH) + (SH >= P ? (m << (SH >= P ? SH - P : 0)) : (m >> (SH < P ? P - SH : 0)));
}

// MODE (forward network only): 0 = radix-2 butterflies on canonical products, 1 = shift form with the forward roots,
// 2 = shift form with the inverse roots (the plain iNTT runs the forward network on inverse roots)
// NAT: the lanes of a contiguous tile are sub-blocks with consecutive bit-reversed indices (tile_nat) instead of columns;
// FULL: every lane of the tile exists (the column count is a multiple of C): no per-lane predicates
template <typename G, bool GS, bool CONTIG, int MODE, int I, bool NAT = false, bool FULL = false>
__device__ __forceinline__ void run_step(const Io<CONTIG>& io, uint64_t* __restrict__ x, const uint64_t* __restrict__ tw, int tid) {
    constexpr int l = G::l, C = G::C, P = G::PADLOG;
    constexpr int NR = Sched<l>::NR;
    constexpr int RHO = GS ? NR - 1 - I : I;
    using rd = Rd<l, RHO>;
    constexpr int K = rd::K, U0 = rd::U0, SH = rd::SH, NE = 1 << K;
    constexpr int LN = (K == 3) ? G::LN3 : G::LN4;
    constexpr int NLG = C / LN;  // lane groups
    constexpr bool FIRST = (I == 0), LAST = (I == NR - 1);
    constexpr int NITEMS = ((1 << l) >> K) * NLG;
    constexpr int NT = G::NT;
#pragma unroll 1
    for (int w = tid; w < NITEMS; w += NT) {
        const int lane0 = (w % NLG) * LN;
        int rest = w / NLG;
        // one padding row per 32: neighbouring threads of the stride-1 radix-16 round must sit 32 rows apart, not 16
        if (SH == 0 && K == 4 && P == 5) rest = (rest & ~3) | ((rest & 1) << 1) | ((rest >> 1) & 1);
        const int j = rest & ((1 << SH) - 1), qh = rest >> SH;
        const int rbase = (qh << (SH + K)) + j;
        uint64_t v[NE][LN];
        uint64_t* xp = x + (size_t)(rbase + (rbase >> P)) * C + lane0;
        if (FIRST) {
            if (!CONTIG) {
                const uint64_t* p = io.in + (size_t)rbase * io.in_row + lane0;
                const size_t step = io.in_row << SH;
#pragma unroll
                for (int m = 0; m < NE; ++m) {
                    if (LN == 2) {
                        const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p + (size_t)m * step);
                        v[m][0] = q.x;
                        v[m][LN - 1] = q.y;
                    } else {
                        v[m][0] = p[(size_t)m * step];
                    }
                }
            } else {
#pragma unroll
                for (int ln = 0; ln < LN; ++ln) {
                    const bool ok = FULL || NAT || lane0 + ln < io.lanes_valid;
                    const uint64_t* p = io.in + (NAT ? io.lane_in[lane0 + ln] : (size_t)(lane0 + ln) * io.in_lane) + rbase;
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
                if (LN == 2) {
                    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(xp + row_off<SH, P>(m) * C);
                    v[m][0] = q.x;
                    v[m][LN - 1] = q.y;
                } else {
                    v[m][0] = xp[row_off<SH, P>(m) * C];
                }
            }
        }
        if constexpr (MODE == 0)
            bfly_regs<K, U0, GS, LN>(v, tw + rd::OFF + qh);
        else
            bfly_shift<K, U0, LN, MODE == 2>(v, tw + rd::OFF + qh + (NAT ? (size_t)lane0 * io.twl : 0), NAT ? io.twl : 0,
                                             (LAST && io.apply_scale) ? &io.scale : nullptr);
        if (LAST) {
            // forward network: a strided pass is never the last one (its outputs stay lazy, no scale), a contiguous pass always
            // is (canonical outputs) -- ntt.cu forward(); the inverse network decides at run time
            if constexpr (GS || CONTIG) {
#pragma unroll
                for (int m = 0; m < NE; ++m) {
#pragma unroll
                    for (int ln = 0; ln < LN; ++ln) {
                        if (MODE == 0 && io.apply_scale)
                            v[m][ln] = gl::mul(v[m][ln], io.scale);  // (the shift form folds the scale into the last round's twiddles)
                        else if (!GS || !io.lazy_out)
                            v[m][ln] = gl::canon_fast(v[m][ln]);
                    }
                }
            }
            if (!CONTIG) {
                uint64_t* p = io.out + (size_t)rbase * io.out_row + lane0;
                const size_t step = io.out_row << SH;
#pragma unroll
                for (int m = 0; m < NE; ++m) {
                    if (LN == 2)
                        *reinterpret_cast<ulonglong2*>(p + (size_t)m * step) = make_ulonglong2(v[m][0], v[m][LN - 1]);
                    else
                        p[(size_t)m * step] = v[m][0];
                }
            } else if constexpr (NAT) {
                // natural-order output: bitrev_l(rbase + (m << SH)) = bitrev_l(rbase) | (bitrev_K(m) << (l - K - SH)); the lanes of
                // a row are neighbours in the output
                const size_t rb = (size_t)(__brev((uint32_t)rbase) >> (32 - l));
#pragma unroll
                for (int ln = 0; ln < LN; ++ln) {
                    uint64_t* p = io.out + (size_t)(lane0 + ln) * io.out_lane + io.br_off;
#pragma unroll
                    for (int m = 0; m < NE; ++m)
                        p[((rb | ((size_t)brev_const((uint32_t)m, K) << (l - K - SH))) << io.br_t) * io.br_mul] = v[m][ln];
                }
            } else {
#pragma unroll
                for (int ln = 0; ln < LN; ++ln) {
                    if (!FULL && lane0 + ln >= io.lanes_valid) continue;
                    uint64_t* p = io.out + (size_t)(lane0 + ln) * io.out_lane + rbase;
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
            for (int m = 0; m < NE; ++m) {
                if (LN == 2)
                    *reinterpret_cast<ulonglong2*>(xp + row_off<SH, P>(m) * C) = make_ulonglong2(v[m][0], v[m][LN - 1]);
                else
                    xp[row_off<SH, P>(m) * C] = v[m][0];
            }
        }
    }
    if (!LAST) __syncthreads();
}

template <typename G, bool GS, bool CONTIG, int MODE, int I, bool NAT = false, bool FULL = false>
__device__ __forceinline__ void run_steps(const Io<CONTIG>& io, uint64_t* x, const uint64_t* tw, int tid) {
    static_assert(!(GS && MODE != 0), "the shift form exists for the forward network only");
    run_step<G, GS, CONTIG, MODE, I, NAT, FULL>(io, x, tw, tid);
    if constexpr (I + 1 < Sched<G::l>::NR) run_steps<G, GS, CONTIG, MODE, I + 1, NAT, FULL>(io, x, tw, tid);
}

// Both kernels are persistent over tiles that share one twiddle table (same sub-block Q, same coset): the table is
// built once per CTA and a.tiles_per_cta tiles stream through it (the per-tile prologue -- a serial chain of squarings
// for the stage constants plus R - 1 products -- was a quarter of the kernel time when paid per tile).

// strided pass: tile = (sub-block Q, C adjacent inner indices) x all 2^l rows.  grid (cosets, sub-blocks x chunks):
// the cosets of one chunk are resident together and walk the same tiles, so the shared input of a coset LDE
// (src_coset_stride == 0) comes from HBM once and from L2 for the other cosets.
template <typename G, bool GS, int MODE = 0>
__global__ void __launch_bounds__(G::NT, G::MINB) tile_strided(const PassArgs a) {
    extern __shared__ __align__(16) uint64_t sm[];
    constexpr int l = G::l, C = G::C, R = 1 << l;
    uint64_t* tw = sm;
    uint64_t* cu = sm + R;
    uint64_t* x = sm + R + 16;
    const int tid = threadIdx.x;
    const uint32_t coset = blockIdx.x;
    const size_t inner = (size_t)1 << (a.M - l);
    constexpr int C_LOG = (C == 8) ? 3 : (C == 4) ? 2 : (C == 2) ? 1 : 0;
    static_assert((1 << C_LOG) == C, "C is 1, 2, 4 or 8");
    const int tps_log = a.M - l - C_LOG;
    const size_t tiles_per_sub = (size_t)1 << tps_log;
    const size_t per_q = a.ncols * tiles_per_sub;  // tiles of one sub-block: (column, tile) pairs
    const size_t chunks_per_q = (per_q + a.tiles_per_cta - 1) / a.tiles_per_cta;
    const uint32_t Q = (uint32_t)(blockIdx.y / chunks_per_q);
    const size_t t0 = (blockIdx.y % chunks_per_q) * a.tiles_per_cta;
    const size_t t1 = t0 + a.tiles_per_cta < per_q ? t0 + a.tiles_per_cta : per_q;
    Io<false> io;
    io.in_row = io.out_row = inner;
    io.in_lane = io.out_lane = 1;
    io.lanes_valid = C;
    io.br_t = -1;
    io.br_mul = io.br_off = 0;
    io.lane_in = nullptr;
    io.twl = 0;
    io.apply_scale = a.apply_scale;
    io.lazy_out = a.lazy_out;
    io.scale = a.scale;
    if (tid == 0) row_constants(a, a.s_last[coset], Q, cu);
    __syncthreads();
    if constexpr (MODE == 0)
        build_twiddles<l>(tw, cu, a.brs, tid, G::NT);
    else
        build_twiddles_pow_round<l, 0>(tw, cu, a.brs, tid, G::NT, a.apply_scale ? a.scale : 1);
    __syncthreads();
    for (size_t t = t0; t < t1; ++t) {
        const size_t col = t >> tps_log, c0 = (t & (tiles_per_sub - 1)) * C;  // tiles_per_sub is a power of two: no 64-bit division per tile
        io.in = a.src + col * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << a.M) + c0;
        io.out = a.dst + col * a.dst_col_stride + coset * a.dst_coset_stride + ((size_t)Q << a.M) + c0;
        run_steps<G, GS, false, MODE, 0>(io, x, tw, tid);
        if (Sched<l>::NR > 1 && t + 1 < t1) __syncthreads();  // the next tile's first round overwrites the staging rows
    }
}

// contiguous pass (M == l): tile = sub-block Q (2^l consecutive elements) of C columns.  grid (sub-blocks, chunks of
// column groups, cosets)
template <typename G, bool GS, int MODE = 0, bool FULL = false>
__global__ void __launch_bounds__(G::NT, G::MINB) tile_contig(const PassArgs a) {
    extern __shared__ __align__(16) uint64_t sm[];
    constexpr int l = G::l, C = G::C, R = 1 << l;
    uint64_t* tw = sm;
    uint64_t* cu = sm + R;
    uint64_t* x = sm + R + 16;
    const int tid = threadIdx.x;
    const uint32_t coset = blockIdx.z;
    const uint32_t Q = blockIdx.x;
    const size_t groups = (a.ncols + C - 1) / C;
    const size_t g0 = (size_t)blockIdx.y * a.tiles_per_cta;
    const size_t g1 = g0 + a.tiles_per_cta < groups ? g0 + a.tiles_per_cta : groups;
    Io<true> io;
    io.in_row = io.out_row = 1;
    io.in_lane = a.src_col_stride;
    io.out_lane = a.dst_col_stride;
    io.br_t = -1;
    io.br_mul = io.br_off = 0;
    io.lane_in = nullptr;
    io.twl = 0;
    io.apply_scale = a.apply_scale;
    io.lazy_out = a.lazy_out;
    io.scale = a.scale;
    if (tid == 0) row_constants(a, a.s_last[coset], Q, cu);
    __syncthreads();
    if constexpr (MODE == 0)
        build_twiddles<l>(tw, cu, a.brs, tid, G::NT);
    else
        build_twiddles_pow_round<l, 0>(tw, cu, a.brs, tid, G::NT, a.apply_scale ? a.scale : 1);
    __syncthreads();
    for (size_t g = g0; g < g1; ++g) {
        const size_t col0 = g * C;
        // the tile a.prefetch groups ahead is pulled into L2 by the copy engine (one bulk prefetch per column: 2^l contiguous
        // elements) while this one is transformed: the first round's loads then wait for L2, not for HBM
        if (a.prefetch && tid < C && g + a.prefetch < g1 && (g + a.prefetch) * C + tid < a.ncols)
            tma::bulk_prefetch_l2(a.src + ((g + a.prefetch) * C + tid) * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << l), (uint32_t)(R * sizeof(uint64_t)));
        io.in = a.src + col0 * a.src_col_stride + coset * a.src_coset_stride + ((size_t)Q << l);
        io.out = a.dst + col0 * a.dst_col_stride + coset * a.dst_coset_stride + ((size_t)Q << l);
        io.lanes_valid = (int)((a.ncols - col0) < (size_t)C ? (a.ncols - col0) : (size_t)C);
        run_steps<G, GS, true, MODE, 0, false, FULL>(io, x, tw, tid);
        if (Sched<l>::NR > 1 && g + 1 < g1) __syncthreads();
    }
}

// last pass with NATURAL-ORDER output (the plain iNTT: coefficients; a natural-order LDE): tile = C sub-blocks x all 2^l
// rows of ONE column, the sub-blocks being those whose bit-reversed indices are consecutive (k0 .. k0 + C - 1): row r of
// sub-block bitrev_t(k) belongs at natural index (bitrev_l(r) << t) + k, so the C lanes of a row are neighbours in the
// output and every store instruction of a warp fills whole 32-byte sectors (the generic bit-reversed store writes 16-byte
// pieces; writing one sub-block per CTA scatters single elements: 2.2x DRAM write traffic, profiles/r02n_*).  The price
// is one twiddle table per lane (the sub-blocks' row constants differ), built once per CTA and amortised over the
// columns it streams.  Shift form only (MODE 1 / 2).  grid (2^t / C, column chunks, cosets)
template <typename G, int MODE>
__global__ void __launch_bounds__(G::NT, G::MINB) tile_nat(const PassArgs a) {
    extern __shared__ __align__(16) uint64_t sm[];
    constexpr int l = G::l, C = G::C, R = 1 << l;
    uint64_t* tw = sm;                                          // [C][R]
    uint64_t* cu = sm + (size_t)C * R;                          // [C][16]
    size_t* lane_in = reinterpret_cast<size_t*>(cu + C * 16);   // [C] (16 slots reserved)
    uint64_t* x = cu + C * 16 + 16;
    const int tid = threadIdx.x;
    const uint32_t coset = blockIdx.z;
    const int t_done = a.L - l;
    const uint32_t k0 = blockIdx.x * C;
    const size_t g0 = (size_t)blockIdx.y * a.tiles_per_cta;
    const size_t g1 = g0 + a.tiles_per_cta < a.ncols ? g0 + a.tiles_per_cta : a.ncols;
    if (tid < C) {
        const uint32_t Q = gl::bitrev32(k0 + tid, t_done);
        row_constants(a, a.s_last[coset], Q, cu + tid * 16);
        lane_in[tid] = (size_t)Q << l;
    }
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < C; ++g) build_twiddles_pow_round<l, 0>(tw + (size_t)g * R, cu + g * 16, a.brs, tid, G::NT, a.apply_scale ? a.scale : 1);
    __syncthreads();
    Io<true> io;
    io.in_row = io.out_row = 1;
    io.in_lane = 0;
    io.out_lane = a.out_mul;
    io.lanes_valid = C;
    io.br_t = t_done;
    io.br_mul = a.out_mul;
    io.br_off = (size_t)k0 * a.out_mul + (a.coset_bits ? gl::bitrev32(coset, a.coset_bits) : 0);
    io.lane_in = lane_in;
    io.twl = R;
    io.apply_scale = a.apply_scale;
    io.lazy_out = a.lazy_out;
    io.scale = a.scale;
    for (size_t col = g0; col < g1; ++col) {
        if (a.prefetch && tid < C && col + a.prefetch < g1)
            tma::bulk_prefetch_l2(a.src + (col + a.prefetch) * a.src_col_stride + coset * a.src_coset_stride + lane_in[tid], (uint32_t)(R * sizeof(uint64_t)));
        io.in = a.src + col * a.src_col_stride + coset * a.src_coset_stride;
        io.out = a.dst + col * a.dst_col_stride;
        run_steps<G, false, true, MODE, 0, true>(io, x, tw, tid);
        if (Sched<l>::NR > 1 && col + 1 < g1) __syncthreads();
    }
}

}  // namespace tile
}  // namespace ntt
}  // namespace ola
