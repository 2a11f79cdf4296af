// Shift-twiddle register rounds of the forward (Cooley-Tukey) network: the same 2^K-row block as bfly_regs<K, U0, false>
// in ntt_tile.cuh, computed as "scale by theta^m, then a 2^K-point transform whose twiddles are powers of two" in the
// 96-bit lazy representation of w96.cuh.
//
// The K stages of a block (stage s pairs row m with m + 2^(K-1-s), twiddle T(s, ql) for the 2^s sub-blocks ql) use
//     T(s, ql) = theta^(2^(K-1-s)) * omega_{2^(s+1)}^{bitrev_s(ql)},   theta = T(K-1, 0)
// (ntt.cu: tw = c_u * BRS[q] with c_{u-1} = c_u^2 and BRS[q] = omega_{2^(u+1)}^{bitrev_u(q)}), so the block maps
// x_0 .. x_{2^K-1} to  sum_m x_m (theta * omega_{2^K}^{bitrev_K(m')})^m  at output row m'.  With y_m = x_m theta^m the
// rest is the network with theta = 1, and omega_{2^(s+1)} = 2^(39 * 2^(5-s)) (mod p) because omega_64 = 2^39 for the
// reference's two-adic generator (goldilocks_field.rs:77; asserted by tests/test_ntt_shift.py and at context creation).
#pragma once
#include "w96.cuh"

namespace ola {
namespace ntt {
namespace tile {

// exponent (mod 192) of the power of two that is omega_{2^(s+1)}^{bitrev_s(ql)}
GL_HD constexpr int shift_expo(int s, int ql) {
    int br = 0;
    for (int i = 0; i < s; ++i) br |= ((ql >> i) & 1) << (s - 1 - i);
    return ((39 << (5 - s)) * br) % 192;
}

GL_HD gl::W96 mul_pow2_sel(gl::W96 v, int e /* 0 < e < 96, a multiple of 12 */) {
    switch (e) {
        case 12: return gl::w96_mul_pow2<12>(v);
        case 24: return gl::w96_mul_pow2<24>(v);
        case 36: return gl::w96_mul_pow2<36>(v);
        case 48: return gl::w96_mul_pow2<48>(v);
        case 60: return gl::w96_mul_pow2<60>(v);
        case 72: return gl::w96_mul_pow2<72>(v);
        default: return gl::w96_mul_pow2<84>(v);
    }
}

// t[(m - 1) << U0] = theta^m for m = 1 .. 2^K - 1 (lane ln reads its own table lane_stride elements further on when the
// lanes are different sub-blocks, tile_nat).  Inputs and outputs are lazy u64 representatives.
// INV: the table holds inverse roots (the plain iNTT runs the forward network on omega^-1): omega^-1 = 2^(192 - 39 j)
template <int K, int U0, int LN, bool INV = false>
// scale0 != nullptr: the table holds theta^m * scale (an output scale folded into the twiddles: 1/n of the iNTT) and
// row 0, which no twiddle multiplies, takes the factor explicitly -- one product per block instead of one per output.
GL_HD void bfly_shift(uint64_t (&v)[1 << K][LN], const uint64_t* __restrict__ t, size_t lane_stride = 0, const uint64_t* scale0 = nullptr) {
    static_assert(K >= 1 && K <= 4, "shift rounds cover up to sixteen rows (omega_16 = 2^156)");
    constexpr int NE = 1 << K;
#pragma unroll
    for (int ln = 0; ln < LN; ++ln) {
        gl::W96 x[NE];
        // every output contains row 0 once: a multiple of p added to it makes all of them non-negative
        x[0] = gl::w96_bias(scale0 ? gl::w96_mul(v[0][ln], *scale0) : gl::w96_from_u64(v[0][ln]));
#pragma unroll
        for (int m = 1; m < NE; ++m) x[m] = gl::w96_mul(v[m][ln], t[ln * lane_stride + ((m - 1) << U0)]);
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const int half = (NE >> 1) >> s;
#pragma unroll
            for (int ql = 0; ql < (1 << s); ++ql) {
                const int e = INV ? (192 - shift_expo(s, ql)) % 192 : shift_expo(s, ql);
#pragma unroll
                for (int jj = 0; jj < half; ++jj) {
                    const int m = (ql << (K - s)) + jj;
                    const gl::W96 p = (e % 96 == 0) ? x[m + half] : mul_pow2_sel(x[m + half], e % 96);
                    const gl::W96 a = x[m];
                    if (e < 96) {
                        x[m] = gl::w96_add(a, p);
                        x[m + half] = gl::w96_sub(a, p);
                    } else {  // 2^(96 + e') = -2^e'
                        x[m] = gl::w96_sub(a, p);
                        x[m + half] = gl::w96_add(a, p);
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < NE; ++m) v[m][ln] = gl::w96_to_u64_nonneg(x[m]);
    }
}

}  // namespace tile
}  // namespace ntt
}  // namespace ola
