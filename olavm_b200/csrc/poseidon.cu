// Poseidon kernels: batched permutation, leaf sponge over row tiles, Merkle level reduction.
//
// Replaces:
//   plonky2/plonky2/src/hash/merkle_tree/mod.rs   new_v2 leaf loop :186-201 (hash_no_pad per row),
//       build_merkle_nodes :311-337 / merkle_tree/concurrent.rs:14-66 (2-to-1 compression per level)
//   plonky2/plonky2/src/hash/hashing.rs           hash_n_to_m_no_pad :84-106, compress :66-74
//
// Leaf hashing reads the LDE COLUMN-major: thread r reads element r of every column, so each warp
// load is one fully coalesced 256-byte segment and no transpose of the LDE matrix (fri/oracle.rs:84)
// is ever materialised.  One thread = one sponge; the permutation is ALU-bound (~500 modular
// multiplications, 64 B of fresh input), so the roofline of these kernels is the integer pipe, not HBM.
#include "common.h"
#include "poseidon.cuh"

#include <cstdlib>

namespace ola {
namespace poseidon {

void init_constants() {
    // canonicalise on upload: add_canonical_u64 (poseidon.rs:482-484) assumes canonical constants
    auto up = [](const void* sym, const uint64_t* src, size_t n) {
        std::vector<uint64_t> t(n);
        for (size_t i = 0; i < n; ++i) t[i] = gl::canon(src[i]);
        OLA_CUDA(cudaMemcpyToSymbol(sym, t.data(), n * sizeof(uint64_t)));
    };
    up(c_round, OLA_ALL_ROUND_CONSTANTS, 360);
    up(c_first, OLA_FAST_PARTIAL_FIRST_ROUND_CONSTANT, 12);
    up(c_partial, OLA_FAST_PARTIAL_ROUND_CONSTANTS, 22);
    up(c_vs, &OLA_FAST_PARTIAL_ROUND_VS[0][0], 22 * 11);
    up(c_whats, &OLA_FAST_PARTIAL_ROUND_W_HATS[0][0], 22 * 11);
    up(c_init, &OLA_FAST_PARTIAL_ROUND_INITIAL_MATRIX[0][0], 11 * 11);
}

__global__ void __launch_bounds__(128, 5) permute_kernel(uint64_t* states, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k] = states[i * 12 + k];
    permute(s);
#pragma unroll
    for (int k = 0; k < 12; ++k) states[i * 12 + k] = s[k];
}

// hash_no_pad over one row; ROWMAJOR: element (r, c) at base[r*ncols + c]; else at base[c*col_stride + r]
// MODE: 0 rolled permutation, 1 fully straight-line (A/B only), 2 straight-line full rounds inside rolled round loops,
// 3 rolled S-box layer + straight-line 22-bit-limb MDS
template <bool ROWMAJOR, int MINB, int MODE = 0>
__global__ void __launch_bounds__(128, MINB) hash_rows_kernel(const uint64_t* __restrict__ base, size_t col_stride, size_t nrows,
                                                       size_t ncols, uint64_t* __restrict__ digests) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k] = 0;
    for (size_t c0 = 0; c0 < ncols; c0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c0 + k < ncols) {
                uint64_t v = ROWMAJOR ? base[r * ncols + c0 + k] : base[(c0 + k) * col_stride + r];
                s[k] = v;
            }
        }
        if (MODE == 1)
            permute_unrolled(s);
        else
            permute<MODE>(s);
    }
    ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + 4 * r);
    d[0] = make_ulonglong2(s[0], s[1]);
    d[1] = make_ulonglong2(s[2], s[3]);
}

// nodes[i] = two_to_one(nodes[2i], nodes[2i+1]) for i in [first, first + count)
template <int MODE>
__global__ void __launch_bounds__(128, 5) merkle_level_kernel(uint64_t* nodes, size_t first, size_t count) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    size_t i = first + k;
    const ulonglong2* ch = reinterpret_cast<const ulonglong2*>(nodes + 8 * i);
    ulonglong2 a = ch[0], b = ch[1], c = ch[2], d = ch[3];
    uint64_t s[12] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, 0, 0, 0, 0};
    permute<MODE>(s);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(nodes + 4 * i);
    o[0] = make_ulonglong2(s[0], s[1]);
    o[1] = make_ulonglong2(s[2], s[3]);
}

// body form of the permutation (OLA_POSEIDON_UNROLLED: 0 rolled, 1 straight-line, 2 straight-line full rounds, 3 limb MDS)
static int poseidon_mode() {
    static const int v = [] {
        const char* e = getenv("OLA_POSEIDON_UNROLLED");
        int m = e ? atoi(e) : 3;  // profiles/r01m_poseidon_sweep.txt: 3 (limb MDS) 116 ms, 0 (rolled) 126 ms, 2 (40 KB of code) 129 ms
        return (m >= 0 && m <= 3) ? m : 3;
    }();
    return v;
}
// OLA_MERKLE_SMALL: levels with at most this many nodes run the straight-line body (0 = never)
static size_t merkle_small_threshold() {
    static const size_t v = [] {
        const char* e = getenv("OLA_MERKLE_SMALL");
        return e ? (size_t)atoll(e) : (size_t)0;
    }();
    return v;
}
void permute_states(ola_ctx* ctx, uint64_t* d_states, size_t n) {
    if (!n) return;
    {
        Launch lz(ctx, "poseidon_permute");
        permute_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_states, n);
    }
    check_launch("permute_kernel");
}
void hash_rows_rowmajor(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, size_t ncols, uint64_t* d_digests) {
    if (!nrows) return;
    {
        Launch lz(ctx, "poseidon_leaves_rowmajor");
        hash_rows_kernel<true, 4, 3><<<(unsigned)((nrows + 127) / 128), 128, 0, ctx->stream>>>(d_rows, 0, nrows, ncols, d_digests);
    }
    check_launch("hash_rows_kernel<row>");
}
void hash_rows_colmajor(ola_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t nrows, size_t ncols,
                        uint64_t* d_digests) {
    if (!nrows) return;
    {
        // resident CTAs per SM (register cap 128/96/80) and body form, tuned on hardware (profiles/)
        static const int minb = [] {
            const char* e = getenv("OLA_POSEIDON_MINB");
            int v = e ? atoi(e) : 5;
            return (v >= 4 && v <= 6) ? v : 5;
        }();
        const int mode = poseidon_mode();
        const unsigned blocks = (unsigned)((nrows + 127) / 128);
        Launch lz(ctx, "poseidon_leaves");
#define OLA_HR(M, U) hash_rows_kernel<false, M, U><<<blocks, 128, 0, ctx->stream>>>(d_cols, col_stride, nrows, ncols, d_digests)
        if (mode == 1) {
            if (minb == 4) OLA_HR(4, 1); else if (minb == 5) OLA_HR(5, 1); else OLA_HR(6, 1);
        } else if (mode == 2) {
            if (minb == 4) OLA_HR(4, 2); else if (minb == 5) OLA_HR(5, 2); else OLA_HR(6, 2);
        } else if (mode == 3) {
            if (minb == 4) OLA_HR(4, 3); else if (minb == 5) OLA_HR(5, 3); else OLA_HR(6, 3);
        } else {
            if (minb == 4) OLA_HR(4, 0); else if (minb == 5) OLA_HR(5, 0); else OLA_HR(6, 0);
        }
#undef OLA_HR
    }
    check_launch("hash_rows_kernel<col>");
}
void merkle_levels(ola_ctx* ctx, uint64_t* d_nodes, size_t nleaves, size_t stop) {
    if (stop < 1) stop = 1;
    for (size_t first = nleaves / 2; first >= stop && first >= 1; first /= 2) {
        {
            Launch lz(ctx, "merkle_level");
            // a level of a few thousand nodes is one permutation deep per thread and latency-bound (one warp per scheduler at
            // most): the straight-line full rounds give that warp twelve independent S-box chains instead of three
            if (first <= merkle_small_threshold())
                merkle_level_kernel<2><<<(unsigned)((first + 127) / 128), 128, 0, ctx->stream>>>(d_nodes, first, first);
            else if (poseidon_mode() == 2)
                merkle_level_kernel<2><<<(unsigned)((first + 127) / 128), 128, 0, ctx->stream>>>(d_nodes, first, first);
            else if (poseidon_mode() == 3)
                merkle_level_kernel<3><<<(unsigned)((first + 127) / 128), 128, 0, ctx->stream>>>(d_nodes, first, first);
            else
                merkle_level_kernel<0><<<(unsigned)((first + 127) / 128), 128, 0, ctx->stream>>>(d_nodes, first, first);
        }
        check_launch("merkle_level_kernel");
        if (first == 1) break;
    }
}

// ---- host permutation for the Fiat-Shamir transcript (iop/challenger.rs:137-152 duplexing) ----
// The transcript is sequential and some callers push millions of elements through it (the Bitwise table's compress challenge
// observes twelve whole columns, generation/builtin.rs:118-131), so the host path uses the same formulation as the kernels:
// MDS rows as 32-bit-half dot products with the small circulant constants, and the fast partial rounds (poseidon.rs:392-421).
static inline uint64_t h_red(unsigned __int128 x) {
    const uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64), EPSV = 0xFFFFFFFFULL, PV = 0xFFFFFFFF00000001ULL;
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & EPSV;
    uint64_t t0;
    const uint64_t borrow = __builtin_sub_overflow(lo, hi_hi, &t0);
    t0 -= EPSV & (0 - borrow);
    const uint64_t t1 = (hi_lo << 32) - hi_lo;
    uint64_t t2;
    const uint64_t carry = __builtin_add_overflow(t0, t1, &t2);
    t2 += EPSV & (0 - carry);
    t2 -= PV & (0 - (uint64_t)(t2 >= PV));
    return t2;
}
static inline uint64_t h_mulmod(uint64_t a, uint64_t b) { return h_red((unsigned __int128)a * b); }
static inline uint64_t h_sbox(uint64_t x) {
    const uint64_t x2 = h_mulmod(x, x), x4 = h_mulmod(x2, x2), x3 = h_mulmod(x, x2);
    return h_mulmod(x3, x4);
}
__attribute__((target_clones("avx2", "default"))) static void h_mds(uint64_t* s) {
    uint32_t C[12], D[12];  // the circulant's first row and the diagonal (all below 2^6)
    for (int i = 0; i < 12; ++i) C[i] = (uint32_t)OLA_MDS_MATRIX_CIRC[i], D[i] = (uint32_t)OLA_MDS_MATRIX_DIAG[i];
    uint32_t lo[24], hi[24];
    for (int i = 0; i < 12; ++i) lo[i] = lo[i + 12] = (uint32_t)s[i], hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32);
    uint64_t al[12], ah[12];
    for (int r = 0; r < 12; ++r) al[r] = ah[r] = 0;
    for (int i = 0; i < 12; ++i)          // output rows are the vector lanes: al[r] += lo[r + i] * C[i]
        for (int r = 0; r < 12; ++r) al[r] += (uint64_t)lo[r + i] * C[i], ah[r] += (uint64_t)hi[r + i] * C[i];
    for (int r = 0; r < 12; ++r) al[r] += (uint64_t)lo[r] * D[r], ah[r] += (uint64_t)hi[r] * D[r];
    for (int r = 0; r < 12; ++r) s[r] = h_red(((unsigned __int128)ah[r] << 32) + al[r]);
}
// sum of up to 2^32 64 x 64-bit products without carry tests: the low and the high words are summed separately (each sum fits 128
// bits), and 2^64 = 2^32 - 1 (mod p) folds the high sum back: total = lo_sum + hi_sum * (2^32 - 1) < 2^101
struct HostAcc {
    unsigned __int128 lo = 0, hi = 0;
    inline void mac(uint64_t a, uint64_t b) {
        const unsigned __int128 p = (unsigned __int128)a * b;
        lo += (uint64_t)p;
        hi += (uint64_t)(p >> 64);
    }
    inline uint64_t reduce() const { return h_red(lo + hi * 0xFFFFFFFFULL); }
};
static inline uint64_t h_add(uint64_t a, uint64_t b) {  // canonical a, b -> canonical a + b, branch-free
    uint64_t s;
    const uint64_t carry = __builtin_add_overflow(a, b, &s);
    s += 0xFFFFFFFFULL & (0 - carry);                                  // + 2^64 mod p
    s -= 0xFFFFFFFF00000001ULL & (0 - (uint64_t)(s >= 0xFFFFFFFF00000001ULL));
    return s;
}
void permute_host(uint64_t s[12]) {
    for (int i = 0; i < 12; ++i) s[i] = gl::canon(s[i]);
    int rc = 0;
    auto full = [&]() {
        for (int r = 0; r < 4; ++r, ++rc) {
            for (int i = 0; i < 12; ++i) s[i] = h_sbox(h_add(s[i], gl::canon(OLA_ALL_ROUND_CONSTANTS[rc * 12 + i])));
            h_mds(s);
        }
    };
    full();
    {  // partial_first_constant_layer + mds_partial_layer_init (poseidon.rs:303-313, :332-358)
        for (int i = 0; i < 12; ++i) s[i] = h_add(s[i], gl::canon(OLA_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]));
        uint64_t q[11];
        for (int c = 0; c < 11; ++c) {
            HostAcc acc;
            for (int r = 1; r < 12; ++r) acc.mac(s[r], OLA_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r - 1][c]);
            q[c] = acc.reduce();
        }
        for (int c = 0; c < 11; ++c) s[c + 1] = q[c];
    }
    for (int r = 0; r < 22; ++r) {  // poseidon.rs:582-588 with mds_partial_layer_fast (:392-421)
        s[0] = h_add(h_sbox(s[0]), gl::canon(OLA_FAST_PARTIAL_ROUND_CONSTANTS[r]));
        HostAcc d;
        d.mac(s[0], OLA_MDS_MATRIX_CIRC[0] + OLA_MDS_MATRIX_DIAG[0]);
        for (int i = 1; i < 12; ++i) d.mac(s[i], OLA_FAST_PARTIAL_ROUND_W_HATS[r][i - 1]);
        const uint64_t s0 = s[0];
        for (int i = 1; i < 12; ++i) s[i] = h_add(s[i], h_mulmod(s0, OLA_FAST_PARTIAL_ROUND_VS[r][i - 1]));
        s[0] = d.reduce();
    }
    rc += 22;
    full();
}



}  // namespace poseidon
}  // namespace ola
