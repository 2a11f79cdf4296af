// Poseidon-Goldilocks permutation (width 12, x^7, 4 + 22 + 4 rounds) as a device function.
//
// Replaces plonky2/plonky2/src/hash/poseidon.rs:593-604 (`Poseidon::poseidon`: full_rounds :566-574,
// partial_rounds :577-590 with the "fast" partial-round decomposition :303-421) and the modes built on
// it in hashing.rs (:66-74 compress, :84-108 overwrite-mode sponge, rate 8).
//
// One thread owns one 12-word state in registers; all round constants come from __constant__ memory
// (every lane of a warp reads the same word in the same cycle -> broadcast).  The MDS layer exploits
// the small circulant coefficients (<= 41): it accumulates 32-bit halves in 64-bit registers and reduces
// once per output, as the reference does with u128 (poseidon.rs:168-189, :236-257).  The dense
// partial-round dot products accumulate 64x64-bit products in a 160-bit register triple and reduce
// once (poseidon.rs:392-409).  Integer pipes only; there is no tensor-core formulation of x^7 S-boxes.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

namespace ola {
namespace poseidon {

#ifdef __CUDACC__
// device copies of the parameter tables (libola_gpu is a single translation unit; filled by init_constants)
static __constant__ uint64_t c_round[360];
static __constant__ uint64_t c_first[12];
static __constant__ uint64_t c_partial[22];
static __constant__ uint64_t c_vs[22 * 11];
static __constant__ uint64_t c_whats[22 * 11];
static __constant__ uint64_t c_init[11 * 11];

// 160-bit accumulator (five 32-bit limbs) for sums of 64x64 products: one carry chain per product
struct acc160 {
    uint32_t w0, w1, w2, w3, w4;
};
__device__ __forceinline__ void acc_zero(acc160& a) { a.w0 = a.w1 = a.w2 = a.w3 = a.w4 = 0; }
__device__ __forceinline__ void acc_mac(acc160& a, uint64_t x, uint64_t y) {
    uint32_t p0, p1, p2, p3;
    gl::mul_limbs(x, y, p0, p1, p2, p3);
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(a.w0), "+r"(a.w1), "+r"(a.w2), "+r"(a.w3), "+r"(a.w4)
        : "r"(p0), "r"(p1), "r"(p2), "r"(p3));
}
// value = (w3:w2:w1:w0) + w4 * 2^128,  2^128 = -2^32 (mod p);  result lazy
__device__ __forceinline__ uint64_t acc_reduce(const acc160& a) {
    uint64_t r = gl::reduce_limbs(a.w0, a.w1, a.w2, a.w3);
    return gl::sub_lc(r, (uint64_t)a.w4 << 32);
}

// x^7 on lazy representatives
__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl::sqr_lazy(x), x4 = gl::sqr_lazy(x2), x3 = gl::mul_lazy(x, x2);
    return gl::mul_lazy(x3, x4);
}

// out[r] = sum_i s[(i+r)%12] * CIRC[i] + s[r]*DIAG[r]   (CIRC = 17,15,41,16,2,28,13,13,39,18,34,20; DIAG[0] = 8)
// lazy in, lazy out
__device__ __forceinline__ void mds_layer(uint64_t* s) {
    constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        uint64_t al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            al += (uint64_t)lo[(i + r) % 12] * C[i];
            ah += (uint64_t)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) {
            al += (uint64_t)lo[0] * 8u;
            ah += (uint64_t)hi[0] * 8u;
        }
        // value = al + ah * 2^32  (al, ah < 2^42): bits above 2^64 fit one limb
        uint64_t l128 = al + (ah << 32);
        uint32_t h96 = (uint32_t)(ah >> 32) + (l128 < al);
        s[r] = gl::reduce96_lazy(l128, h96);
    }
}

__device__ __forceinline__ void full_round_unrolled(uint64_t* s, int round_ctr) {
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = sbox7(gl::add_lc(s[i], c_round[round_ctr * 12 + i]));
    mds_layer(s);
}

// ---- code-size-bounded form ----
// The fully unrolled permutation is ~6000 SASS instructions (~100 KB); warps of the resident CTAs drift through
// it independently, so the SM's instruction working set is the whole body and the measured dominant stall was
// "no instruction" (profiles/poseidon_r01f_ncu_summary.md).  Below, the 12-lane S-box layer and the circulant MDS
// are rolled into short loops over a ROTATING register file (all indices inside a loop body are static; the
// rotation is register moves), the 11x11 initial matrix is one row per iteration, and both groups of four full
// rounds share one body, so the whole permutation fits the instruction cache.
__device__ __forceinline__ void full_round(uint64_t* s, int round_ctr) {
    // S-box layer: 4 iterations x 3 lanes; rotating left by 3 each time returns the state to natural order
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int k = round_ctr * 12 + it * 3;
        const uint64_t a = sbox7(gl::add_lc(s[0], c_round[k]));
        const uint64_t b = sbox7(gl::add_lc(s[1], c_round[k + 1]));
        const uint64_t c = sbox7(gl::add_lc(s[2], c_round[k + 2]));
#pragma unroll
        for (int j = 0; j < 9; ++j) s[j] = s[j + 3];
        s[9] = a;
        s[10] = b;
        s[11] = c;
    }
    // MDS layer: 3 iterations x 4 rows of the circulant; rotating the operand halves by 4 turns rows 4k..4k+3 into
    // rows 0..3, and the outputs queue up in o
    constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
    uint64_t o[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
        o[i] = 0;
    }
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
        const uint32_t diag = it == 0 ? 8u : 0u;  // MDS_MATRIX_DIAG[0] = 8 applies to row 0 only
        uint64_t nw[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint64_t al = 0, ah = 0;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                al += (uint64_t)lo[(i + r) % 12] * C[i];
                ah += (uint64_t)hi[(i + r) % 12] * C[i];
            }
            if (r == 0) {
                al += (uint64_t)lo[0] * diag;
                ah += (uint64_t)hi[0] * diag;
            }
            const uint64_t l128 = al + (ah << 32);
            const uint32_t h96 = (uint32_t)(ah >> 32) + (l128 < al);
            nw[r] = gl::reduce96_lazy(l128, h96);
        }
        uint32_t tl[4], th[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tl[j] = lo[j];
            th[j] = hi[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            lo[j] = lo[j + 4];
            hi[j] = hi[j + 4];
            o[j] = o[j + 4];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lo[8 + j] = tl[j];
            hi[8 + j] = th[j];
            o[8 + j] = nw[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = o[i];
}

// MDS layer on 22-bit limbs, straight-line.  Every circulant coefficient is <= 41 and their sum (with the diagonal 8)
// is 264 < 2^9, so a 22-bit limb times the coefficients, summed over a row, stays below 2^31: the 3 x 144 products are
// plain 32-bit IMADs (2 cycles of the fmaheavy pipe each) instead of 2 x 144 IMAD.WIDE (4 cycles each), and nothing
// has to rotate.  The limb split and the recombination are shifts and adds on the ALU pipe, which has the headroom
// (profiles/r01m_poseidon_pipe_model.md).  lazy in, lazy out.
__device__ __forceinline__ void mds_layer_limbs(uint64_t* s) {
    constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        l0[i] = (uint32_t)s[i] & 0x3FFFFFu;
        l1[i] = (uint32_t)(s[i] >> 22) & 0x3FFFFFu;
        l2[i] = (uint32_t)(s[i] >> 44);
    }
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const uint32_t c = C[i] + ((r == 0 && i == 0) ? 8u : 0u);  // MDS_MATRIX_DIAG[0] = 8
            a0 += l0[(i + r) % 12] * c;
            a1 += l1[(i + r) % 12] * c;
            a2 += l2[(i + r) % 12] * c;
        }
        // value = a0 + a1 * 2^22 + a2 * 2^44  (a0, a1 < 2^31, a2 < 2^29): 73 bits
        const uint64_t t = (uint64_t)a0 + ((uint64_t)a1 << 22);
        const uint64_t lo = t + ((uint64_t)a2 << 44);
        const uint32_t x2 = (a2 >> 20) + (lo < t);
        s[r] = gl::reduce96_lazy(lo, x2);
    }
}

// rolled S-box layer (as full_round) + straight-line limb MDS
__device__ __forceinline__ void full_round_limbs(uint64_t* s, int round_ctr) {
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int k = round_ctr * 12 + it * 3;
        const uint64_t a = sbox7(gl::add_lc(s[0], c_round[k]));
        const uint64_t b = sbox7(gl::add_lc(s[1], c_round[k + 1]));
        const uint64_t c = sbox7(gl::add_lc(s[2], c_round[k + 2]));
#pragma unroll
        for (int j = 0; j < 9; ++j) s[j] = s[j + 3];
        s[9] = a;
        s[10] = b;
        s[11] = c;
    }
    mds_layer_limbs(s);
}

// partial_first_constant_layer + mds_partial_layer_init (poseidon.rs:303-313, :332-358), one output column per
// iteration; results queue up in q
__device__ __forceinline__ void partial_init(uint64_t* s) {
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = gl::add_lc(s[i], c_first[i]);
    uint64_t q[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) q[j] = 0;
#pragma unroll 1
    for (int c = 0; c < 11; ++c) {
        acc160 acc;
        acc_zero(acc);
#pragma unroll
        for (int r = 1; r < 12; ++r) acc_mac(acc, s[r], c_init[(r - 1) * 11 + c]);
        const uint64_t v = acc_reduce(acc);
#pragma unroll
        for (int j = 0; j < 10; ++j) q[j] = q[j + 1];
        q[10] = v;
    }
#pragma unroll
    for (int c = 0; c < 11; ++c) s[c + 1] = q[c];
}

// 22 partial rounds (poseidon.rs:582-588, mds_partial_layer_fast :392-421)
__device__ __forceinline__ void partial_rounds(uint64_t* s) {
#pragma unroll 1
    for (int r = 0; r < 22; ++r) {
        s[0] = gl::add_lc(sbox7(s[0]), c_partial[r]);
        acc160 d;
        {  // d = s[0] * 25  (MDS_MATRIX_CIRC[0] + MDS_MATRIX_DIAG[0]): two IMAD.WIDE, a 70-bit value
            const uint64_t lo = (uint64_t)(uint32_t)s[0] * 25u, hi = (s[0] >> 32) * 25u;
            const uint64_t mid = (lo >> 32) + (uint32_t)hi;  // < 2^33
            d.w0 = (uint32_t)lo;
            d.w1 = (uint32_t)mid;
            d.w2 = (uint32_t)(hi >> 32) + (uint32_t)(mid >> 32);  // hi < 2^37: no overflow
            d.w3 = d.w4 = 0;
        }
#pragma unroll
        for (int i = 1; i < 12; ++i) acc_mac(d, s[i], c_whats[r * 11 + i - 1]);
        uint64_t s0 = s[0];
#pragma unroll
        for (int i = 1; i < 12; ++i) s[i] = gl::mad_lazy(s0, c_vs[r * 11 + i - 1], s[i]);
        s[0] = acc_reduce(d);
    }
}

// s: any u64 representatives in, CANONICAL representatives out.
// FORM 0: rolled S-box layer and rolled circulant MDS on 32-bit halves; 2: straight-line full rounds (A/B only: 40 KB
// of code, slower); 3: rolled S-box layer + straight-line MDS on 22-bit limbs.
template <int FORM = 3>
__device__ __forceinline__ void permute(uint64_t* s) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            if (FORM == 2)
                full_round_unrolled(s, half * 26 + r);
            else if (FORM == 3)
                full_round_limbs(s, half * 26 + r);
            else
                full_round(s, half * 26 + r);
        }
        if (half == 0) {
            partial_init(s);
            partial_rounds(s);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = gl::canon_fast(s[i]);
}

// the straight-line form (kept for A/B measurements: OLA_POSEIDON_UNROLLED=1)
__device__ __forceinline__ void permute_unrolled(uint64_t* s) {
#pragma unroll 1
    for (int r = 0; r < 4; ++r) full_round_unrolled(s, r);
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = gl::add_lc(s[i], c_first[i]);
    {
        acc160 acc[11];
#pragma unroll
        for (int c = 0; c < 11; ++c) acc_zero(acc[c]);
#pragma unroll
        for (int r = 1; r < 12; ++r) {
#pragma unroll
            for (int c = 0; c < 11; ++c) acc_mac(acc[c], s[r], c_init[(r - 1) * 11 + c]);
        }
#pragma unroll
        for (int c = 0; c < 11; ++c) s[c + 1] = acc_reduce(acc[c]);
    }
    partial_rounds(s);
#pragma unroll 1
    for (int r = 0; r < 4; ++r) full_round_unrolled(s, 26 + r);
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = gl::canon_fast(s[i]);
}
#endif  // __CUDACC__

// ---- host launchers (poseidon.cu) ----
}  // namespace poseidon
}  // namespace ola

struct ola_ctx;
namespace ola {
namespace poseidon {
void init_constants();
// states [n][12] in place
void permute_states(ola_ctx* ctx, uint64_t* d_states, size_t n);
// row-major leaves [nrows][ncols] -> digests [nrows][4]
void hash_rows_rowmajor(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, size_t ncols, uint64_t* d_digests);
// column-major leaves: element (row r, column c) at d_cols[c*col_stride + r]; digests [nrows][4]
void hash_rows_colmajor(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t nrows, size_t ncols,
                        uint64_t* d_digests);
// heap-ordered tree: d_nodes is [2*nleaves][4] with the leaf digests already at [nleaves, 2*nleaves);
// fills nodes [stop, nleaves) level by level (stop >= 1; stop = 2^cap_height fills down to the cap level)
void merkle_levels(ola_ctx* ctx, uint64_t* d_nodes, size_t nleaves, size_t stop);
// host-side permutation (transcript / challenger only)
void permute_host(uint64_t state[12]);
}  // namespace poseidon
}  // namespace ola
