// Multi-table STARK prover on the device.
//
// Replaces:
//   circuits/src/stark/prover.rs:79-327      prove_with_traces (orchestration, transcript order)
//   circuits/src/stark/prover.rs:330-567     prove_single_table
//   circuits/src/stark/prover.rs:571-705     compute_quotient_polys   -> quotient_kernel<Air>
//   circuits/src/stark/cross_table_lookup.rs:224-310  cross_table_lookup_data / partial_products -> ctl_values + scan
//   circuits/src/stark/permutation.rs:103-160 compute_permutation_z_polys -> perm_values + exclusive scan
//   circuits/src/stark/proof.rs:199-246      StarkOpeningSet::new -> eval_partial_kernel + host combine
//   circuits/src/stark/vanishing_poly.rs:20-47, permutation.rs:302-360, cross_table_lookup.rs:380-419
//
// Everything large stays in HBM between phases (trace values, the three PolynomialBatches of a table, quotient
// values); the host sees caps, openings and query rows only.  LDE rows are read column-major in leaf order:
// a warp's 32 rows are 32 consecutive u64 of every column (coalesced), and the "next row" of leaf position p is
// bitrev(bitrev(p)+1): for every warp but a 2^-(L-5) fraction that is again 32 consecutive u64.
#include "stark.h"

#include <algorithm>

#include "air/registry.cuh"
#include "batch.h"
#include "fri.h"
#include "ntt.h"
#include "tma.cuh"
#include "verify.h"

namespace ola {
namespace nccl {
bool allgather_side(ola_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream);
bool has_side_comm(const ola_ctx* ctx);
}  // namespace nccl
// batch.cu
__global__ void canon_copy_kernel(uint64_t* dst, const uint64_t* src, size_t n);
namespace stark {

using air::Consumer;
using air::Fp;
using air::Row;

struct DevBuf {
    uint64_t* p = nullptr;
    DevBuf() {}
    explicit DevBuf(size_t n) { dev_alloc(&p, n); }
    ~DevBuf() {
        if (p) ola::dev_free(p);
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// ---- device descriptors of linear-combination columns / CTL instances / permutation instances ----------------
struct DevLc {
    int off, cnt;
    uint64_t constant;
    int single;  // >= 0: the combination is exactly that trace column (Column::single) -- no multiply
    int pad_;
};
struct DevCtl {
    int col_off, col_cnt;  // into lcs[]
    int filter;            // index into lcs[] or -1
    int twin;              // the instance over the same columns with the next challenge pair, or -1
    int is_twin;           // this instance is evaluated together with an earlier one (its `twin`)
    int pad_;
    uint64_t beta, gamma;
};
struct DevPermInst {
    int pair_off, pair_cnt;  // into perm_pairs[] (lhs,rhs interleaved)
    uint64_t beta, gamma;
};
struct DevPermBatch {
    int inst_off, inst_cnt;
};
// The quotient kernel's view of the CTL section: one DevSide per lookup side of this table (a TableWithColumns,
// cross_table_lookup.rs:173-194) with its two challenge instances.  GrandProductChallenge::combine is sum_q beta^q col_q +
// gamma and a Column is sum_t coef_t trace[col_t] + const, so an instance's combination is ONE dot product over trace
// columns: the host folds beta^q coef_t into per-term weights (w0 / w1 for the two challenges) and the constants into c0 / c1.
struct DevTerm {
    uint64_t off;  // byte offset of the trace column inside the LDE buffer (column * column stride * 8)
    uint64_t w0, w1;
};
struct DevSide {
    int term_off, term_cnt;  // into terms[]
    int filt_off, filt_cnt;  // into fterms[] (w0 = coefficient); filt_cnt < 0: no filter (the filter is 1)
    uint64_t c0, c1;         // sum_q beta_j^q const_q + gamma_j
    uint64_t filt_const;
    int z0, z1;              // Z columns of the two instances (index into zs, permutation Zs included); z1 < 0: none
    int k0, k1;              // index of each instance's first-row constraint among the table's constraints (transition: + 1)
    int filt_single, pad_;   // >= 0: the filter is exactly that trace column
};
struct DevTables {  // all arrays live in one device allocation
    const DevSide* sides;
    const DevTerm* terms;
    const DevTerm* fterms;
    int nsides, nterms, nfterms;
    const DevLc* lcs;
    const int* lc_col;
    const uint64_t* lc_coef;
    const DevCtl* ctls;
    int nctl;
    const DevPermInst* perm_insts;
    const DevPermBatch* perm_batches;
    const int* perm_pairs;
    int nperm_batches;
};

__device__ __forceinline__ Fp eval_lc(const DevTables& d, int k, const uint64_t* __restrict__ base, size_t stride, size_t r) {
    const DevLc lc = d.lcs[k];
    if (lc.single >= 0) return Fp(__ldg(base + (size_t)lc.single * stride + r));
    Fp s(0);
    for (int i = 0; i < lc.cnt; ++i) {
        const Fp v(__ldg(base + (size_t)d.lc_col[lc.off + i] * stride + r));
        const uint64_t k = d.lc_coef[lc.off + i];
        s += k == 1 ? v : v * Fp(k);
    }
    return s + Fp(lc.constant);
}
// GrandProductChallenge::combine (permutation.rs:61-72): reduce_with_powers(terms, beta) + gamma
__device__ __forceinline__ Fp ctl_combine(const DevTables& d, const DevCtl& c, const uint64_t* __restrict__ base, size_t stride, size_t r) {
    Fp s(0), beta(c.beta);
    for (int i = c.col_cnt - 1; i >= 0; --i) s = s * beta + eval_lc(d, c.col_off + i, base, stride, r);
    return s + Fp(c.gamma);
}

// ---- CTL Z columns: per-row factor, then an inclusive prefix product (partial_products, cross_table_lookup.rs:284-310)
__global__ void ctl_values_kernel(DevTables d, const uint64_t* __restrict__ trace, size_t n, uint64_t* __restrict__ zs /* [nctl][n] */, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevCtl c = d.ctls[blockIdx.y];
    uint64_t filter = c.filter >= 0 ? eval_lc(d, c.filter, trace, n, i).v : 1;
    uint64_t v = 1;
    if (filter == 1)
        v = ctl_combine(d, c, trace, n, i).v;
    else if (filter != 0)
        atomicOr(err, 1);  // "Non-binary filter?" (cross_table_lookup.rs:305)
    zs[(size_t)blockIdx.y * n + i] = v;
}

// permutation Z factors: prod_inst lhs / prod_inst rhs  (compute_permutation_z_poly, permutation.rs:129-160)
// Each thread handles PV_ROWS rows a grid-stride apart (coalesced) and shares ONE field inversion between them
// (Montgomery's trick: 3 (PV_ROWS - 1) multiplications instead of PV_ROWS - 1 further ~95-multiplication inversions).
static constexpr int PV_ROWS = 4;
__global__ void perm_values_kernel(DevTables d, const uint64_t* __restrict__ trace, size_t n, uint64_t* __restrict__ zs) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const DevPermBatch b = d.perm_batches[blockIdx.y];
    uint64_t num[PV_ROWS], den[PV_ROWS];
#pragma unroll
    for (int r = 0; r < PV_ROWS; ++r) {
        const size_t i = i0 + (size_t)r * stride;
        Fp nm(1), dn(1);
        if (i < n) {
            for (int k = 0; k < b.inst_cnt; ++k) {
                const DevPermInst in = d.perm_insts[b.inst_off + k];
                Fp l(in.gamma), rr(in.gamma), w(1), beta(in.beta);
                for (int p = 0; p < in.pair_cnt; ++p) {
                    l += Fp(trace[(size_t)d.perm_pairs[2 * (in.pair_off + p)] * n + i]) * w;
                    rr += Fp(trace[(size_t)d.perm_pairs[2 * (in.pair_off + p) + 1] * n + i]) * w;
                    w *= beta;
                }
                nm *= l;
                dn *= rr;
            }
        }
        num[r] = nm.v;
        den[r] = dn.v;
    }
    // prefix products of the denominators, one inversion, then peel the inverses off from the back
    uint64_t pre[PV_ROWS];
    pre[0] = den[0];
#pragma unroll
    for (int r = 1; r < PV_ROWS; ++r) pre[r] = gl::mul(pre[r - 1], den[r]);
    uint64_t inv = gl::inv(pre[PV_ROWS - 1]);
#pragma unroll
    for (int r = PV_ROWS - 1; r >= 0; --r) {
        const uint64_t inv_r = r ? gl::mul(inv, pre[r - 1]) : inv;
        if (r) inv = gl::mul(inv, den[r]);
        const size_t i = i0 + (size_t)r * stride;
        if (i < n) zs[(size_t)blockIdx.y * n + i] = gl::mul(num[r], inv_r);
    }
}

// ---- prefix products over columns [ncols][n], in place.  exclusive: out[i] = prod_{j<i}; else prod_{j<=i}
static constexpr int SC_RUN = 8, SC_THREADS = 256, SC_CHUNK = SC_RUN * SC_THREADS;
__global__ void __launch_bounds__(SC_THREADS) scan_chunk_kernel(const uint64_t* __restrict__ data, size_t n, uint64_t* __restrict__ chunkp, size_t nchunks) {
    __shared__ uint64_t sh[SC_THREADS];
    const uint64_t* c = data + (size_t)blockIdx.y * n;
    size_t start = (size_t)blockIdx.x * SC_CHUNK + (size_t)threadIdx.x * SC_RUN;
    uint64_t acc = 1;
    for (int k = 0; k < SC_RUN; ++k)
        if (start + k < n) acc = gl::mul(acc, c[start + k]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SC_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = gl::mul(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) chunkp[(size_t)blockIdx.y * nchunks + blockIdx.x] = sh[0];
}
__global__ void scan_carry_kernel(uint64_t* chunkp, size_t nchunks, size_t ncols) {
    size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    uint64_t* p = chunkp + col * nchunks;
    uint64_t acc = 1;
    for (size_t s = 0; s < nchunks; ++s) {
        uint64_t v = p[s];
        p[s] = acc;  // exclusive prefix of chunk products
        acc = gl::mul(acc, v);
    }
}
__global__ void __launch_bounds__(SC_THREADS) scan_emit_kernel(uint64_t* __restrict__ data, size_t n, const uint64_t* __restrict__ chunkp, size_t nchunks, int exclusive) {
    __shared__ uint64_t sh[SC_THREADS];
    uint64_t* c = data + (size_t)blockIdx.y * n;
    const int t = threadIdx.x;
    size_t start = (size_t)blockIdx.x * SC_CHUNK + (size_t)t * SC_RUN;
    uint64_t v[SC_RUN];
    uint64_t acc = 1;
    for (int k = 0; k < SC_RUN; ++k) {
        v[k] = start + k < n ? c[start + k] : 1;
        acc = gl::mul(acc, v[k]);
    }
    sh[t] = acc;
    __syncthreads();
    for (int s = 1; s < SC_THREADS; s <<= 1) {  // inclusive Hillis-Steele scan of the run products
        uint64_t x = sh[t];
        if (t >= s) x = gl::mul(x, sh[t - s]);
        __syncthreads();
        sh[t] = x;
        __syncthreads();
    }
    uint64_t carry = chunkp[(size_t)blockIdx.y * nchunks + blockIdx.x];
    if (t > 0) carry = gl::mul(carry, sh[t - 1]);
    for (int k = 0; k < SC_RUN; ++k) {
        if (start + k >= n) break;
        if (exclusive) {
            c[start + k] = carry;
            carry = gl::mul(carry, v[k]);
        } else {
            carry = gl::mul(carry, v[k]);
            c[start + k] = carry;
        }
    }
}

// ---- quotient: one thread per LDE point of the size n*2^qdb quotient domain ------------------------------------
struct QuotArgs {
    const uint64_t* trace_lde;  // [cols][L]
    const uint64_t* zs_lde;     // [nzs][L]
    size_t L;                   // LDE size (column stride)
    int log_n, qdb;
    uint64_t alpha0, alpha1;
    uint64_t g, g_inv, n_field;  // subgroup generator, its inverse, n as a field element
    uint64_t zh[8], zh_inv[8];   // Z_H on the 2^qdb cosets, indexed by (i mod 2^qdb)
    const uint64_t* pw;          // omega_{2^32} power tables (forward)
    uint64_t* out;               // [2][n << qdb]
    DevTables d;
    int num_perm_zs;
    uint64_t compress_challenge;  // Bitwise / Program only (canonical)
    const uint64_t* weights;      // [K][2]: alpha_0^(K-1-k), alpha_1^(K-1-k) for the table's K constraints (AIR, permutation, CTL)
    int nweights;                 // K
    // coset shard: this rank evaluates points [r_offset, r_offset + npoints) of the quotient domain; its LDE buffers
    // start at leaf r_offset (row r of the domain is local row r - r_offset); out[j][r - r_offset], j-stride out_stride
    size_t r_offset, npoints, out_stride;
};

__device__ __forceinline__ uint64_t pow_omega_fwd(const uint64_t* __restrict__ pw, uint32_t E) {
    uint64_t r = __ldg(pw + 4096 + (E >> 22));
    r = gl::mul(r, __ldg(pw + 2048 + ((E >> 11) & 2047u)));
    return gl::mul(r, __ldg(pw + (E & 2047u)));
}

// Shared-memory image of what every thread of a CTA reads uniformly: the constraint weights and the CTL descriptors.
// One elected thread stages it with 1-D bulk copies (TMA engine, tma.cuh) while the others compute their row addresses.
struct QuotSmem {
    static __host__ __device__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
    static __host__ __device__ size_t weights_off() { return 16; }  // [0,8): mbarrier
    static __host__ __device__ size_t sides_off(int K) { return weights_off() + align16((size_t)K * 16); }
    static __host__ __device__ size_t terms_off(int K, int nsides) { return sides_off(K) + align16((size_t)nsides * sizeof(DevSide)); }
    static __host__ __device__ size_t fterms_off(int K, int nsides, int nterms) { return terms_off(K, nsides) + align16((size_t)nterms * sizeof(DevTerm)); }
    static __host__ __device__ size_t bytes(int K, int nsides, int nterms, int nfterms) {
        return fterms_off(K, nsides, nterms) + align16((size_t)nfterms * sizeof(DevTerm));
    }
};

template <class Air, int MINB = 3>
__global__ void __launch_bounds__(128, MINB) quotient_kernel(const QuotArgs a) {
    extern __shared__ __align__(16) unsigned char q_smem[];
    const DevTables& d = a.d;
    uint64_t* bar = reinterpret_cast<uint64_t*>(q_smem);
    const uint64_t* s_w = reinterpret_cast<const uint64_t*>(q_smem + QuotSmem::weights_off());
    const DevSide* s_sides = reinterpret_cast<const DevSide*>(q_smem + QuotSmem::sides_off(a.nweights));
    const DevTerm* s_terms = reinterpret_cast<const DevTerm*>(q_smem + QuotSmem::terms_off(a.nweights, d.nsides));
    const DevTerm* s_fterms = reinterpret_cast<const DevTerm*>(q_smem + QuotSmem::fterms_off(a.nweights, d.nsides, d.nterms));
    if (threadIdx.x == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_barrier_init();
        const uint32_t b_w = (uint32_t)QuotSmem::align16((size_t)a.nweights * 16), b_s = (uint32_t)QuotSmem::align16((size_t)d.nsides * sizeof(DevSide)),
                       b_t = (uint32_t)QuotSmem::align16((size_t)d.nterms * sizeof(DevTerm)), b_f = (uint32_t)QuotSmem::align16((size_t)d.nfterms * sizeof(DevTerm));
        tma::mbar_arrive_expect_tx(bar, b_w + b_s + b_t + b_f);
        tma::bulk_g2s((void*)s_w, a.weights, b_w, bar);
        if (b_s) tma::bulk_g2s((void*)s_sides, d.sides, b_s, bar);
        if (b_t) tma::bulk_g2s((void*)s_terms, d.terms, b_t, bar);
        if (b_f) tma::bulk_g2s((void*)s_fterms, d.fterms, b_f, bar);
    }
    // Row mapping.  Warp w of the launch handles the 32 rows whose natural indices are j = lane' * (n/32) + m for the block
    // counter m = w mod (n/32): consecutive warps hold consecutive m, and the NEXT rows of block m are the rows of block
    // m + 1 -- so a CTA's next-row loads are its neighbour warp's local-row loads (L1), and the CTA after it reuses the
    // last block through L2.  In leaf order those 32 rows are the consecutive positions (bitrev(m) << 5 | lane).
    const size_t r_lin = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = r_lin < a.npoints;
    const uint32_t nmask = (1u << a.log_n) - 1;
    uint32_t pos_lin = (uint32_t)r_lin & nmask;
    if (a.log_n > 5) {
        const uint32_t m = pos_lin >> 5, lane = pos_lin & 31u;
        pos_lin = (gl::bitrev32(m, a.log_n - 5) << 5) | lane;
    }
    const size_t r_raw = (r_lin & ~(size_t)nmask) | pos_lin;  // local row (leaf order) in this rank's buffers
    const size_t rg = r_raw + a.r_offset;                      // index in the quotient domain (leaf order)
    const uint32_t coset = (uint32_t)(rg >> a.log_n), pos = (uint32_t)rg & nmask;
    const uint32_t j = gl::bitrev32(pos, a.log_n);
    const uint32_t pos_next = gl::bitrev32((j + 1) & nmask, a.log_n);
    // local rows in this rank's LDE buffers (the next row stays inside the coset)
    const size_t r = r_raw, r_next = (((size_t)coset << a.log_n) | pos_next) - a.r_offset;
    // natural LDE index k = bitrev_{log_n+3}(r) = j*8 + bitrev3(coset); x = 7 * omega_{8n}^k
    const uint32_t cb = gl::bitrev32(coset, 3);
    const uint32_t kk = (j << 3) | cb;
    const int lde_bits = a.log_n + 3;
    const uint32_t E = lde_bits >= 32 ? kk : (kk << (32 - lde_bits));
    const uint64_t x = gl::mul(gl::GEN, pow_omega_fwd(a.pw, E));
    const uint32_t zi = cb >> (3 - a.qdb);  // (k / step) mod 2^qdb
    // L_0(x) = Z_H(x) / (n (x - 1)),  L_last(x) = Z_H(x) / (n (g x - 1))   (verifier.rs:380-396); x is never in H
    const uint64_t d0 = gl::mul(a.n_field, gl::sub(x, 1)), d1 = gl::mul(a.n_field, gl::sub(gl::mul(a.g, x), 1));
    const uint64_t inv01 = gl::inv(gl::mul(d0, d1));
    tma::mbar_wait(bar, 0);
    if (!active) return;
    Consumer yc;
    yc.acc0.clear();
    yc.acc1.clear();
    yc.tr0.clear();
    yc.tr1.clear();
    yc.w = s_w;
    yc.k = 0;
    yc.z_last = Fp(gl::sub(x, a.g_inv));
    yc.lagrange_first = Fp(gl::mul(a.zh[zi], gl::mul(inv01, d1)));
    yc.lagrange_last = Fp(gl::mul(a.zh[zi], gl::mul(inv01, d0)));
    const Row lv{a.trace_lde, a.L, r}, nv{a.trace_lde, a.L, r_next};
    Air::eval(lv, nv, yc, Fp(a.compress_challenge));

    // eval_permutation_checks (permutation.rs:302-360)
    if (a.num_perm_zs > 0) {
        for (int i = 0; i < a.num_perm_zs; ++i) yc.constraint_first_row(Fp(a.zs_lde[(size_t)i * a.L + r]) - air::one());
        for (int i = 0; i < d.nperm_batches; ++i) {
            const DevPermBatch b = d.perm_batches[i];
            Fp lhs(1), rhs(1);
            for (int q = 0; q < b.inst_cnt; ++q) {
                const DevPermInst in = d.perm_insts[b.inst_off + q];
                Fp l(0), rr(0), beta(in.beta);
                for (int p = in.pair_cnt - 1; p >= 0; --p) {  // ReducingFactor::reduce_ext
                    l = l * beta + lv[d.perm_pairs[2 * (in.pair_off + p)]];
                    rr = rr * beta + lv[d.perm_pairs[2 * (in.pair_off + p) + 1]];
                }
                lhs *= l + Fp(in.gamma);
                rhs *= rr + Fp(in.gamma);
            }
            Fp zl(a.zs_lde[(size_t)i * a.L + r]), zn(a.zs_lde[(size_t)i * a.L + r_next]);
            yc.constraint(zn * rhs - zl * lhs);
        }
    }
    // eval_cross_table_lookup_checks (cross_table_lookup.rs:380-419): per instance a first-row constraint
    // (local_z - select(filter, combined)) and a transition constraint (next_z - local_z select(next filter, next
    // combined)), at positions k and k + 1 of the consumer's sequence.  Because every constraint carries its explicit
    // weight the order of evaluation is free: the two challenge instances of a side share every trace load, their
    // combinations are dot products accumulated unreduced, and the transition constraints of all instances are summed
    // (weighted) before the single multiplication by z_last.
    const char* __restrict__ tr_r = reinterpret_cast<const char*>(a.trace_lde + r);
    const char* __restrict__ tr_n = reinterpret_cast<const char*>(a.trace_lde + r_next);
    for (int s = 0; s < d.nsides; ++s) {
        const DevSide sd = s_sides[s];
        air::Wide l0, l1, n0, n1;
        l0.clear();
        l1.clear();
        n0.clear();
        n1.clear();
#pragma unroll 2
        for (int t = 0; t < sd.term_cnt; ++t) {
            const DevTerm tm = s_terms[sd.term_off + t];
            const uint64_t v = __ldg(reinterpret_cast<const uint64_t*>(tr_r + tm.off)), vn = __ldg(reinterpret_cast<const uint64_t*>(tr_n + tm.off));
            l0.mac(v, tm.w0);
            l1.mac(v, tm.w1);
            n0.mac(vn, tm.w0);
            n1.mac(vn, tm.w1);
        }
        Fp lf(1), nf(1);
        if (sd.filt_single >= 0) {
            const uint64_t* col = a.trace_lde + (size_t)sd.filt_single * a.L;
            lf = Fp(__ldg(col + r));
            nf = Fp(__ldg(col + r_next));
        } else if (sd.filt_cnt >= 0) {
            air::Wide fl, fn;
            fl.clear();
            fn.clear();
            for (int t = 0; t < sd.filt_cnt; ++t) {
                const DevTerm tm = s_fterms[sd.filt_off + t];
                fl.mac(__ldg(reinterpret_cast<const uint64_t*>(tr_r + tm.off)), tm.w0);
                fn.mac(__ldg(reinterpret_cast<const uint64_t*>(tr_n + tm.off)), tm.w0);
            }
            lf = Fp(fl.reduce()) + Fp(sd.filt_const);
            nf = Fp(fn.reduce()) + Fp(sd.filt_const);
        }
        const bool filtered = sd.filt_cnt >= 0;
        const Fp one_lf = air::one() - lf, one_nf = air::one() - nf;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int z = h == 0 ? sd.z0 : sd.z1;
            if (z < 0) break;
            Fp lc = Fp((h == 0 ? l0 : l1).reduce()) + Fp(h == 0 ? sd.c0 : sd.c1);
            Fp nc = Fp((h == 0 ? n0 : n1).reduce()) + Fp(h == 0 ? sd.c0 : sd.c1);
            if (filtered) {  // select(filter, x) = filter * x + 1 - filter
                lc = lf * lc + one_lf;
                nc = nf * nc + one_nf;
            }
            const uint64_t* zcol = a.zs_lde + (size_t)z * a.L;
            const Fp local_z(__ldg(zcol + r)), next_z(__ldg(zcol + r_next));
            const Fp c_first = (local_z - lc) * yc.lagrange_first;
            const Fp c_trans = next_z - local_z * nc;
            const int k = h == 0 ? sd.k0 : sd.k1;
            const ulonglong2 wf = *reinterpret_cast<const ulonglong2*>(s_w + 2 * k);
            const ulonglong2 wt = *reinterpret_cast<const ulonglong2*>(s_w + 2 * k + 2);
            yc.acc0.mac(c_first.v, wf.x);
            yc.acc1.mac(c_first.v, wf.y);
            yc.tr0.mac(c_trans.v, wt.x);
            yc.tr1.mac(c_trans.v, wt.y);
        }
    }
    Fp acc0, acc1;
    yc.finish(acc0, acc1);
    a.out[r] = gl::mul(acc0.v, a.zh_inv[zi]);
    a.out[a.out_stride + r] = gl::mul(acc1.v, a.zh_inv[zi]);
}

static void launch_quotient(ola_ctx* ctx, int table_id, QuotArgs a) {
    const size_t size = a.npoints;
    if (size == 0) return;
    // 128 threads x 3 CTAs/SM: the CPU table's body needs ~168 registers; larger blocks and block-wide barriers pacing the
    // instruction stream both measured slower (profiles/quotient_r01h_summary.md)
    const int threads = 128;
    const unsigned blocks = (unsigned)((size + threads - 1) / threads);
    const size_t smem = QuotSmem::bytes(a.nweights, a.d.nsides, a.d.nterms, a.d.nfterms);
    OLA_CHECK(smem <= 48 * 1024, OLA_ERR_INTERNAL, "quotient descriptors exceed the default shared-memory window");
    Launch lz(ctx, "quotient");
    switch (table_id) {
        case T_CPU: {
            // resident CTAs per SM for the widest body (register cap 255 / 168 / 128 / 102 / 80), tuned on hardware:
            // profiles/r01m_quotient_minb_sweep.txt (2^20 rows: 89.9 / 64.1 / 51.0 ms for 2 / 3 / 4)
            static const int minb = [] { const char* e = getenv("OLA_QUOT_MINB"); return e ? atoi(e) : 4; }();
            if (minb == 4)
                quotient_kernel<air::Cpu, 4><<<blocks, threads, smem, ctx->stream>>>(a);
            else if (minb == 5)
                quotient_kernel<air::Cpu, 5><<<blocks, threads, smem, ctx->stream>>>(a);
            else if (minb == 6)
                quotient_kernel<air::Cpu, 6><<<blocks, threads, smem, ctx->stream>>>(a);
            else if (minb == 2)
                quotient_kernel<air::Cpu, 2><<<blocks, threads, smem, ctx->stream>>>(a);
            else
                quotient_kernel<air::Cpu, 3><<<blocks, threads, smem, ctx->stream>>>(a);
            break;
        }
        case T_MEMORY: quotient_kernel<air::Memory><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_CMP: quotient_kernel<air::Cmp><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_RANGECHECK: quotient_kernel<air::RangeCheck><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_BITWISE: quotient_kernel<air::Bitwise><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_POSEIDON: quotient_kernel<air::Poseidon><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_POSEIDON_CHUNK: quotient_kernel<air::PoseidonChunk><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_STORAGE: quotient_kernel<air::StorageAccess><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_TAPE: quotient_kernel<air::Tape><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_SCCALL: quotient_kernel<air::SCCall><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_PROGRAM: quotient_kernel<air::Program><<<blocks, threads, smem, ctx->stream>>>(a); break;
        case T_PROG_CHUNK: quotient_kernel<air::ProgChunk><<<blocks, threads, smem, ctx->stream>>>(a); break;
        default: throw Error(OLA_ERR_INVALID_ARG, "no constraint kernel for table " + std::to_string(table_id));
    }
}

// any non-zero in data[first, last) of each column -> flag
__global__ void nonzero_kernel(const uint64_t* __restrict__ data, size_t col_stride, size_t first, size_t last, int* flag) {
    size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= last) return;
    if (data[(size_t)blockIdx.y * col_stride + i] != 0) atomicOr(flag, 1);
}

// ---- openings: partial Horner sums of every coefficient column at an extension point -------------------------
static constexpr int EV_RUN = 16, EV_THREADS = 256, EV_CHUNK = EV_RUN * EV_THREADS;
struct EvParams {
    gl::ext2 z;
    gl::ext2 zk[EV_RUN];  // z^k, k < RUN
    gl::ext2 zr[8];       // z^(RUN * 2^k)
};
// A thread's run is a dot product of RUN base-field coefficients with the powers z^k: accumulated unreduced (air::Wide, two
// multiply-accumulates per coefficient) instead of RUN dependent extension-field Horner steps.
__global__ void __launch_bounds__(EV_THREADS) eval_partial_kernel(const uint64_t* __restrict__ coeffs, size_t n, EvParams p, uint64_t* __restrict__ partial /* [ncols][nchunks][2] */,
                                                               size_t nchunks) {
    __shared__ gl::ext2 sh[EV_THREADS];
    const uint64_t* c = coeffs + (size_t)blockIdx.y * n;
    const int t = threadIdx.x;
    size_t start = (size_t)blockIdx.x * EV_CHUNK + (size_t)t * EV_RUN;
    air::Wide a0, a1;
    a0.clear();
    a1.clear();
#pragma unroll
    for (int k = 0; k < EV_RUN; ++k) {
        const size_t j = start + k;
        const uint64_t v = j < n ? __ldg(c + j) : 0;
        a0.mac(v, p.zk[k].c0);
        a1.mac(v, p.zk[k].c1);
    }
    sh[t] = gl::make2(a0.reduce(), a1.reduce());
    __syncthreads();
    for (int k = 0; (1 << k) < EV_THREADS; ++k) {
        gl::ext2 v = sh[t];
        if ((t & ((2 << k) - 1)) == 0) v = gl::add(v, gl::mul(p.zr[k], sh[t + (1 << k)]));
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    if (t == 0) {
        size_t o = ((size_t)blockIdx.y * nchunks + blockIdx.x) * 2;
        partial[o] = sh[0].c0;
        partial[o + 1] = sh[0].c1;
    }
}

static std::vector<E> eval_batch_at(ola_ctx* ctx, const ola_batch* b, size_t first_col, size_t ncols, E z) {
    const size_t n = (size_t)1 << b->log_n;
    std::vector<E> out;
    if (!ncols) return out;
    const size_t nchunks = (n + EV_CHUNK - 1) / EV_CHUNK;
    EvParams p;
    p.z = z;
    {
        E zk = gl::make2(1, 0);
        for (int k = 0; k < EV_RUN; ++k) {
            p.zk[k] = zk;
            zk = gl::mul(zk, z);
        }
    }
    E zr = gl::pow(z, (uint64_t)EV_RUN);
    for (int k = 0; k < 8; ++k) {
        p.zr[k] = zr;
        zr = gl::mul(zr, zr);
    }
    // multi-GPU: the coefficients are replicated, so each rank evaluates a slice of the columns and the results are summed
    // (zeros elsewhere); single GPU: the slice is everything
    const size_t per = (ncols + ctx->world - 1) / ctx->world;
    const size_t c_lo = std::min(ncols, (size_t)ctx->rank * per), c_hi = std::min(ncols, c_lo + per), mine = c_hi - c_lo;
    std::vector<uint64_t> res(2 * ncols, 0);
    if (mine) {
        uint64_t* d_part = ctx_scratch(ctx, mine * nchunks * 2);
        {
            Launch lz(ctx, "openings_eval");
            eval_partial_kernel<<<dim3((unsigned)nchunks, (unsigned)mine), EV_THREADS, 0, ctx->stream>>>(b->d_coeffs + (first_col + c_lo) * n, n, p, d_part, nchunks);
        }
        check_launch("eval_partial_kernel");
        std::vector<uint64_t> h(mine * nchunks * 2);
        OLA_CUDA(cudaMemcpyAsync(h.data(), d_part, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        const E zc = gl::pow(z, (uint64_t)EV_CHUNK);
        for (size_t c = 0; c < mine; ++c) {
            E acc = gl::make2(0, 0);
            for (size_t s = nchunks; s-- > 0;) acc = gl::add(gl::mul(acc, zc), gl::make2(h[(c * nchunks + s) * 2], h[(c * nchunks + s) * 2 + 1]));
            res[2 * (c_lo + c)] = gl::canon(acc.c0);
            res[2 * (c_lo + c) + 1] = gl::canon(acc.c1);
        }
    }
    if (ctx->world > 1) {
        DevBuf d_res(2 * ncols);
        OLA_CUDA(cudaMemcpyAsync(d_res.p, res.data(), res.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        comm_allreduce(ctx, d_res.p, 2 * ncols);
        OLA_CUDA(cudaMemcpyAsync(res.data(), d_res.p, res.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    for (size_t c = 0; c < ncols; ++c) out.push_back(gl::make2(res[2 * c], res[2 * c + 1]));
    return out;
}

// ---- host-side descriptor builder --------------------------------------------------------------------------
struct CtlInstance {  // one Z column of a table (CtlZData, cross_table_lookup.rs:196-202)
    Challenge ch;
    const TableWithColumns* twc;
};
struct DescBuilder {
    std::vector<DevLc> lcs;
    std::vector<int> lc_col;
    std::vector<uint64_t> lc_coef;
    std::vector<DevCtl> ctls;
    std::vector<DevPermInst> insts;
    std::vector<DevPermBatch> batches;
    std::vector<int> pairs;
    std::vector<const TableWithColumns*> ctl_sides;
    int add_lc(const Column& c) {
        DevLc l;
        l.off = (int)lc_col.size();
        l.cnt = (int)c.lc.size();
        l.constant = gl::canon(c.constant);
        l.pad_ = 0;
        l.single = (c.lc.size() == 1 && gl::canon(c.lc[0].second) == 1 && l.constant == 0) ? c.lc[0].first : -1;
        for (auto& t : c.lc) {
            lc_col.push_back(t.first);
            lc_coef.push_back(gl::canon(t.second));
        }
        lcs.push_back(l);
        return (int)lcs.size() - 1;
    }
    void add_ctl(const CtlInstance& ci) {
        DevCtl c;
        c.col_off = (int)lcs.size();
        c.col_cnt = (int)ci.twc->columns.size();
        for (auto& col : ci.twc->columns) add_lc(col);
        c.filter = ci.twc->has_filter ? add_lc(ci.twc->filter) : -1;
        c.beta = ci.ch.beta;
        c.gamma = ci.ch.gamma;
        c.twin = -1;
        c.is_twin = 0;
        c.pad_ = 0;
        for (size_t k = 0; k < ctls.size(); ++k)
            if (ctl_sides[k] == ci.twc && ctls[k].twin < 0 && !ctls[k].is_twin) {
                ctls[k].twin = (int)ctls.size();
                c.is_twin = 1;
                break;
            }
        ctl_sides.push_back(ci.twc);
        ctls.push_back(c);
    }
    // the quotient kernel's flattened CTL section (DevSide / DevTerm); base_k = number of constraints in front of it
    std::vector<DevSide> sides;
    std::vector<DevTerm> terms, fterms;
    void build_sides(int base_k, int num_perm_zs, uint64_t col_stride_bytes) {
        for (size_t i = 0; i < ctls.size(); ++i) {
            const DevCtl& c = ctls[i];
            if (c.is_twin) continue;
            const int j = c.twin;
            DevSide sd;
            memset(&sd, 0, sizeof(sd));
            sd.term_off = (int)terms.size();
            F b0 = 1, b1 = 1, c0 = 0, c1 = 0;
            for (int q = 0; q < c.col_cnt; ++q) {
                const DevLc& lc = lcs[c.col_off + q];
                for (int t = 0; t < lc.cnt; ++t) {
                    DevTerm tm;
                    tm.off = (uint64_t)lc_col[lc.off + t] * col_stride_bytes;
                    tm.w0 = gl::mul(b0, lc_coef[lc.off + t]);
                    tm.w1 = gl::mul(b1, lc_coef[lc.off + t]);
                    terms.push_back(tm);
                }
                c0 = gl::add(c0, gl::mul(b0, lc.constant));
                c1 = gl::add(c1, gl::mul(b1, lc.constant));
                b0 = gl::mul(b0, c.beta);
                if (j >= 0) b1 = gl::mul(b1, ctls[j].beta);
            }
            sd.term_cnt = (int)terms.size() - sd.term_off;
            sd.c0 = gl::add(c0, c.gamma);
            sd.c1 = j >= 0 ? gl::add(c1, ctls[j].gamma) : 0;
            sd.filt_single = -1;
            if (c.filter >= 0) {
                const DevLc& lc = lcs[c.filter];
                if (lc.single >= 0) {
                    sd.filt_single = lc.single;
                    sd.filt_off = 0;
                    sd.filt_cnt = 0;
                } else {
                    sd.filt_off = (int)fterms.size();
                    for (int t = 0; t < lc.cnt; ++t) fterms.push_back(DevTerm{(uint64_t)lc_col[lc.off + t] * col_stride_bytes, lc_coef[lc.off + t], 0});
                    sd.filt_cnt = lc.cnt;
                    sd.filt_const = lc.constant;
                }
            } else {
                sd.filt_cnt = -1;
            }
            sd.z0 = num_perm_zs + (int)i;
            sd.z1 = j >= 0 ? num_perm_zs + j : -1;
            sd.k0 = base_k + 2 * (int)i;
            sd.k1 = j >= 0 ? base_k + 2 * j : -1;
            sides.push_back(sd);
        }
    }
};
struct DevDesc {
    uint64_t* mem = nullptr;
    DevTables t{};
    ~DevDesc() {
        if (mem) ola::dev_free(mem);
    }
    void upload(ola_ctx* ctx, const DescBuilder& b) {
        // every array starts on a 16-byte boundary and is followed by slack: the quotient kernel stages the side / term
        // arrays with bulk copies whose size is rounded up to 16 bytes
        auto a8 = [](size_t bytes) { return ((bytes + 15) / 16) * 2; };
        size_t o_lcs = 0, o_col = o_lcs + a8(b.lcs.size() * sizeof(DevLc)), o_coef = o_col + a8(b.lc_col.size() * 4),
               o_ctl = o_coef + a8(b.lc_coef.size() * 8), o_inst = o_ctl + a8(b.ctls.size() * sizeof(DevCtl)),
               o_bat = o_inst + a8(b.insts.size() * sizeof(DevPermInst)), o_pair = o_bat + a8(b.batches.size() * sizeof(DevPermBatch)),
               o_side = o_pair + a8(b.pairs.size() * 4), o_term = o_side + a8(b.sides.size() * sizeof(DevSide)),
               o_fterm = o_term + a8(b.terms.size() * sizeof(DevTerm)), total = o_fterm + a8(b.fterms.size() * sizeof(DevTerm)) + 2;
        std::vector<uint64_t> h(total, 0);
        auto put = [&](size_t off, const void* src, size_t bytes) {
            if (bytes) memcpy(&h[off], src, bytes);
        };
        put(o_lcs, b.lcs.data(), b.lcs.size() * sizeof(DevLc));
        put(o_col, b.lc_col.data(), b.lc_col.size() * 4);
        put(o_coef, b.lc_coef.data(), b.lc_coef.size() * 8);
        put(o_ctl, b.ctls.data(), b.ctls.size() * sizeof(DevCtl));
        put(o_inst, b.insts.data(), b.insts.size() * sizeof(DevPermInst));
        put(o_bat, b.batches.data(), b.batches.size() * sizeof(DevPermBatch));
        put(o_pair, b.pairs.data(), b.pairs.size() * 4);
        put(o_side, b.sides.data(), b.sides.size() * sizeof(DevSide));
        put(o_term, b.terms.data(), b.terms.size() * sizeof(DevTerm));
        put(o_fterm, b.fterms.data(), b.fterms.size() * sizeof(DevTerm));
        dev_alloc(&mem, total);
        OLA_CUDA(cudaMemcpyAsync(mem, h.data(), total * 8, cudaMemcpyHostToDevice, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        t.lcs = (const DevLc*)(mem + o_lcs);
        t.lc_col = (const int*)(mem + o_col);
        t.lc_coef = mem + o_coef;
        t.ctls = (const DevCtl*)(mem + o_ctl);
        t.nctl = (int)b.ctls.size();
        t.perm_insts = (const DevPermInst*)(mem + o_inst);
        t.perm_batches = (const DevPermBatch*)(mem + o_bat);
        t.perm_pairs = (const int*)(mem + o_pair);
        t.nperm_batches = (int)b.batches.size();
        t.sides = (const DevSide*)(mem + o_side);
        t.terms = (const DevTerm*)(mem + o_term);
        t.fterms = (const DevTerm*)(mem + o_fterm);
        t.nsides = (int)b.sides.size();
        t.nterms = (int)b.terms.size();
        t.nfterms = (int)b.fterms.size();
    }
};

static void prefix_products(ola_ctx* ctx, uint64_t* d_data, size_t ncols, size_t n, bool exclusive) {
    if (!ncols) return;
    const size_t nchunks = (n + SC_CHUNK - 1) / SC_CHUNK;
    uint64_t* d_chunk = ctx_scratch(ctx, ncols * nchunks);
    {
        Launch lz(ctx, "scan_chunks");
        scan_chunk_kernel<<<dim3((unsigned)nchunks, (unsigned)ncols), SC_THREADS, 0, ctx->stream>>>(d_data, n, d_chunk, nchunks);
    }
    {
        Launch lz(ctx, "scan_carry");
        scan_carry_kernel<<<(unsigned)((ncols + 63) / 64), 64, 0, ctx->stream>>>(d_chunk, nchunks, ncols);
    }
    {
        Launch lz(ctx, "scan_emit");
        scan_emit_kernel<<<dim3((unsigned)nchunks, (unsigned)ncols), SC_THREADS, 0, ctx->stream>>>(d_data, n, d_chunk, nchunks, exclusive ? 1 : 0);
    }
    check_launch("prefix_products");
}

struct BatchHolder {  // RAII for ola_batch
    ola_batch* b = nullptr;
    BatchHolder() {}
    ~BatchHolder() { reset(); }
    void reset() {
        if (b) {
            batch_release(b);
            delete b;
            b = nullptr;
        }
    }
    BatchHolder(const BatchHolder&) = delete;
    BatchHolder& operator=(const BatchHolder&) = delete;
};

// the full Merkle cap of a commitment; for a coset shard: one all-gather of this rank's cap entries
static Cap batch_cap(ola_ctx* ctx, const ola_batch* b) {
    Cap c((size_t)1 << Config::cap_height);
    if (b->shard_bits == b->rate_bits) {
        batch_get_cap(ctx, b, (uint64_t*)c.data());
        return c;
    }
    const size_t nloc = (size_t)1 << b->local_cap_height();
    OLA_CHECK(nloc * (size_t)ctx->world == c.size(), OLA_ERR_INTERNAL, "cap shard size");
    DevBuf d_full(c.size() * 4);
    comm_allgather(ctx, b->d_nodes + 4 * nloc, d_full.p, nloc * 32);
    OLA_CUDA(cudaMemcpyAsync(c.data(), d_full.p, c.size() * 32, cudaMemcpyDeviceToHost, ctx->stream));
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    return c;
}
// PolynomialBatch commitment of this proof: the whole LDE on one GPU, this rank's cosets under ola_set_comm
static ola_batch* commit(ola_ctx* ctx, const uint64_t* d_cols, size_t ncols, uint32_t log_n, bool is_coeffs) {
    if (ctx->world == 1) return batch_commit(ctx, d_cols, true, ncols, log_n, is_coeffs, Config::rate_bits, Config::cap_height);
    return batch_commit_dist(ctx, d_cols, ncols, log_n, is_coeffs, Config::rate_bits, Config::cap_height);
}

// prove_single_table (prover.rs:330-567)
static StarkProof prove_single_table(ola_ctx* ctx, const TableInfo& t, const Config& cfg, const uint64_t* d_trace, uint32_t degree_bits,
                                     const ola_batch* trace_commit, const std::vector<CtlInstance>& ctl, Challenger& ch) {
    const size_t n = (size_t)1 << degree_bits;
    const std::vector<uint32_t> arities = fri_arities(degree_bits);
    uint32_t total_ar = 0;
    for (auto a : arities) total_ar += a;
    OLA_CHECK(total_ar <= degree_bits + Config::rate_bits - Config::cap_height, OLA_ERR_INVALID_ARG, "FRI total reduction arity is too large.");
    ch.at(STAGE_TABLE_BEGIN);
    ch.compact();

    // permutation challenges and instances (get_n_grand_product_challenge_sets, get_permutation_batches)
    DescBuilder db;
    std::vector<std::vector<Challenge>> perm_sets;
    if (!t.permutation_pairs.empty()) {
        for (int s = 0; s < t.permutation_batch_size(); ++s) {
            std::vector<Challenge> set;
            for (uint32_t k = 0; k < Config::num_challenges; ++k) {
                F b = ch.get_challenge();
                F g = ch.get_challenge();
                set.push_back({b, g});
            }
            perm_sets.push_back(set);
        }
        std::vector<std::pair<const PermutationPair*, int>> all;
        for (auto& p : t.permutation_pairs)
            for (uint32_t c = 0; c < Config::num_challenges; ++c) all.push_back({&p, (int)c});
        const size_t bs = (size_t)t.permutation_batch_size();
        for (size_t s = 0; s < all.size(); s += bs) {
            DevPermBatch b;
            b.inst_off = (int)db.insts.size();
            b.inst_cnt = 0;
            for (size_t i = 0; i < bs && s + i < all.size(); ++i) {
                DevPermInst in;
                in.pair_off = (int)db.pairs.size() / 2;
                in.pair_cnt = (int)all[s + i].first->column_pairs.size();
                for (auto& cp : all[s + i].first->column_pairs) {
                    db.pairs.push_back(cp.first);
                    db.pairs.push_back(cp.second);
                }
                in.beta = perm_sets[i][all[s + i].second].beta;
                in.gamma = perm_sets[i][all[s + i].second].gamma;
                db.insts.push_back(in);
                b.inst_cnt++;
            }
            db.batches.push_back(b);
        }
    }
    for (auto& ci : ctl) db.add_ctl(ci);
    const size_t num_perm_zs = db.batches.size();
    // the table's constraints in consumer order: the AIR's, then per permutation Z a first-row check and per batch a
    // transition check, then two per CTL instance (vanishing_poly.rs:20-47)
    const int k_air = verify::air_constraint_count(t);
    const int k_total = k_air + 2 * (int)num_perm_zs + 2 * (int)ctl.size();
    db.build_sides(k_air + 2 * (int)num_perm_zs, (int)num_perm_zs, ((uint64_t)8 << trace_commit->leaf_bits()));
    const size_t nzs = num_perm_zs + ctl.size();
    OLA_CHECK(nzs > 0, OLA_ERR_INVALID_ARG, "No CTL?");
    DevDesc desc;
    desc.upload(ctx, db);

    // Z columns: values then prefix products
    DevBuf d_zs(nzs * n), d_err(1);
    OLA_CUDA(cudaMemsetAsync(d_err.p, 0, 8, ctx->stream));
    if (num_perm_zs) {
        {
            Launch lz(ctx, "perm_values");
            perm_values_kernel<<<dim3((unsigned)((n + 128 * PV_ROWS - 1) / (128 * PV_ROWS)), (unsigned)num_perm_zs), 128, 0, ctx->stream>>>(desc.t, d_trace, n, d_zs.p);
        }
        check_launch("perm_values_kernel");
        prefix_products(ctx, d_zs.p, num_perm_zs, n, true);
    }
    if (!ctl.empty()) {
        {
            Launch lz(ctx, "ctl_values");
            ctl_values_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)ctl.size()), 128, 0, ctx->stream>>>(desc.t, d_trace, n, d_zs.p + num_perm_zs * n, (int*)d_err.p);
        }
        check_launch("ctl_values_kernel");
        prefix_products(ctx, d_zs.p + num_perm_zs * n, ctl.size(), n, false);
        int herr = 0;
        OLA_CUDA(cudaMemcpyAsync(&herr, d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        OLA_CHECK(herr == 0, OLA_ERR_INVALID_ARG, "Non-binary filter?");
    }
    BatchHolder zs_commit;
    zs_commit.b = commit(ctx, d_zs.p, nzs, degree_bits, false);
    StarkProof proof;
    proof.trace_cap = batch_cap(ctx, trace_commit);
    proof.zs_cap = batch_cap(ctx, zs_commit.b);
    ch.at(STAGE_ZS_CAP);
    ch.observe_cap(proof.zs_cap);
    ch.at(STAGE_ALPHAS);
    const std::vector<F> alphas = ch.get_challenges(2);
    const F alpha0 = alphas[0], alpha1 = alphas[1];

    // ---- compute_quotient_polys
    const int qdf = t.quotient_degree_factor();
    int qdb = 0;
    while ((1 << qdb) < qdf) qdb++;
    OLA_CHECK((uint32_t)qdb <= Config::rate_bits, OLA_ERR_INVALID_ARG, "Having constraints of degree higher than the rate is not supported yet.");
    const size_t qsize = n << qdb;
    // Coset shard: rank r owns leaf blocks [r*per, (r+1)*per) of every LDE; the quotient domain is blocks [0, 2^qdb).
    // Each rank evaluates the blocks it owns (none for high ranks of a low-degree table), then the values are
    // all-gathered block-wise; qstride = distance between the two alpha columns of d_q.
    const int per = (1 << Config::rate_bits) / ctx->world, blk_lo = ctx->rank * per;
    const int qblocks = std::max(0, std::min(per, (1 << qdb) - blk_lo));
    const size_t qstride = ctx->world == 1 ? qsize : (n << Config::rate_bits);
    DevBuf d_q(2 * qstride);
    DevBuf d_qsend(ctx->world == 1 ? 1 : 2 * (size_t)per * n);
    {
        QuotArgs a;
        memset(&a, 0, sizeof(a));
        a.trace_lde = trace_commit->d_lde;
        a.zs_lde = zs_commit.b->d_lde;
        a.L = (size_t)1 << trace_commit->leaf_bits();  // column stride of this rank's LDE buffers
        a.r_offset = (size_t)blk_lo * n;
        a.npoints = (size_t)qblocks * n;
        a.out_stride = ctx->world == 1 ? qsize : (size_t)per * n;
        a.log_n = (int)degree_bits;
        a.qdb = qdb;
        a.alpha0 = alpha0;
        a.alpha1 = alpha1;
        a.g = gl::root_of_unity((int)degree_bits);
        a.g_inv = gl::inv(a.g);
        a.n_field = (uint64_t)n % gl::P;
        const F g_pow_n = gl::pow(gl::GEN, (uint64_t)n);  // ZeroPolyOnCoset::new (zero_poly_coset.rs:19-33)
        F w = gl::root_of_unity(qdb), xx = 1;
        for (int i = 0; i < (1 << qdb); ++i) {
            a.zh[i] = gl::sub(gl::mul(g_pow_n, xx), 1);
            a.zh_inv[i] = gl::inv(a.zh[i]);
            xx = gl::mul(xx, w);
        }
        a.pw = ctx->tw.pw[0];
        a.out = ctx->world == 1 ? d_q.p : d_qsend.p;
        a.d = desc.t;
        a.num_perm_zs = (int)num_perm_zs;
        a.compress_challenge = t.compress_challenge;
        // weights of the consumer's Horner accumulation: constraint k of K carries alpha_j^(K-1-k)
        std::vector<uint64_t> h_w(2 * (size_t)k_total + 2, 0);
        for (int j = 0; j < 2; ++j) {
            const F al = j == 0 ? alpha0 : alpha1;
            F pw = 1;
            for (int k = k_total - 1; k >= 0; --k) {
                h_w[2 * (size_t)k + j] = pw;
                pw = gl::mul(pw, al);
            }
        }
        DevBuf d_w(h_w.size());
        OLA_CUDA(cudaMemcpyAsync(d_w.p, h_w.data(), h_w.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        a.weights = d_w.p;
        a.nweights = k_total;
        launch_quotient(ctx, t.id, a);
        check_launch("quotient_kernel");
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));  // h_w / d_w go out of scope
    }
    if (ctx->world > 1)
        for (int j = 0; j < 2; ++j) comm_allgather(ctx, d_qsend.p + (size_t)j * per * n, d_q.p + (size_t)j * qstride, (size_t)per * n * 8);
    // values at bit-reversed positions on 7*H_{n 2^qdb} -> natural coefficients (coset_ifft, prover.rs:700-704)
    ntt::inverse_from_leaf_order(ctx, d_q.p, qstride, 2, (int)degree_bits + qdb, gl::GEN);
    if (cfg.check_quotient_degree && (size_t)qdf * n < qsize) {
        OLA_CUDA(cudaMemsetAsync(d_err.p, 0, 8, ctx->stream));
        size_t cnt = qsize - (size_t)qdf * n;
        nonzero_kernel<<<dim3((unsigned)((cnt + 255) / 256), 2), 256, 0, ctx->stream>>>(d_q.p, qstride, (size_t)qdf * n, qsize, (int*)d_err.p);
        count_launch(ctx);
        int herr = 0;
        OLA_CUDA(cudaMemcpyAsync(&herr, d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        if (herr) throw Error(OLA_ERR_QUOTIENT_DEGREE, std::string(t.name) + ": Quotient has failed, the vanishing polynomial is not divisible by Z_H");
    }
    // all_quotient_chunks: [alpha][chunk] (prover.rs:463-478)
    DevBuf d_chunks((size_t)2 * qdf * n);
    for (int j = 0; j < 2; ++j)
        OLA_CUDA(cudaMemcpyAsync(d_chunks.p + (size_t)j * qdf * n, d_q.p + (size_t)j * qstride, (size_t)qdf * n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    BatchHolder q_commit;
    q_commit.b = commit(ctx, d_chunks.p, (size_t)2 * qdf, degree_bits, true);
    proof.quotient_cap = batch_cap(ctx, q_commit.b);
    ch.at(STAGE_QUOTIENT_CAP);
    ch.observe_cap(proof.quotient_cap);

    ch.at(STAGE_ZETA);
    const E zeta = ch.get_ext();
    const F g = gl::root_of_unity((int)degree_bits);
    {
        E zp = zeta;
        for (uint32_t k = 0; k < degree_bits; ++k) zp = gl::sqr(zp);
        if (gl::eq(zp, gl::make2(1, 0))) throw Error(OLA_ERR_ZETA_IN_SUBGROUP, "Opening point is in the subgroup.");
    }
    // ---- StarkOpeningSet::new
    const E zeta_next = gl::mul(zeta, g);
    const F g_inv = gl::inv(g);
    OpeningSet& os = proof.openings;
    os.local_values = eval_batch_at(ctx, trace_commit, 0, t.columns, zeta);
    os.next_values = eval_batch_at(ctx, trace_commit, 0, t.columns, zeta_next);
    os.zs = eval_batch_at(ctx, zs_commit.b, 0, nzs, zeta);
    os.zs_next = eval_batch_at(ctx, zs_commit.b, 0, nzs, zeta_next);
    for (auto& e : eval_batch_at(ctx, zs_commit.b, num_perm_zs, nzs - num_perm_zs, gl::make2(g_inv, 0))) os.ctl_zs_last.push_back(e.c0);
    os.quotient = eval_batch_at(ctx, q_commit.b, 0, (size_t)2 * qdf, zeta);
    // observe_openings(to_fri_openings) (proof.rs:248-283)
    ch.at(STAGE_OPENINGS);
    for (auto& v : os.local_values) ch.observe_ext(v);
    for (auto& v : os.zs) ch.observe_ext(v);
    for (auto& v : os.quotient) ch.observe_ext(v);
    for (auto& v : os.next_values) ch.observe_ext(v);
    for (auto& v : os.zs_next) ch.observe_ext(v);
    for (auto& v : os.ctl_zs_last) ch.observe_ext(gl::make2(v, 0));
    // ---- fri_instance (stark.rs:87-150) and opening proof
    fri::Instance inst;
    fri::BatchInfo b0, b1, b2;
    b0.point = zeta;
    for (int c = 0; c < t.columns; ++c) b0.polys.push_back({0, c});
    for (size_t c = 0; c < nzs; ++c) b0.polys.push_back({1, (int)c});
    for (int c = 0; c < 2 * qdf; ++c) b0.polys.push_back({2, c});
    b1.point = zeta_next;
    for (int c = 0; c < t.columns; ++c) b1.polys.push_back({0, c});
    for (size_t c = 0; c < nzs; ++c) b1.polys.push_back({1, (int)c});
    b2.point = gl::make2(g_inv, 0);
    for (size_t c = num_perm_zs; c < nzs; ++c) b2.polys.push_back({1, (int)c});
    inst.batches = {b0, b1, b2};
    proof.fri = fri::prove_openings(ctx, inst, {trace_commit, zs_commit.b, q_commit.b}, ch, degree_bits);
    return proof;
}

// prove_with_traces (prover.rs:79-327) + Buffer::write_all_proof (serialization.rs:377-393)
std::vector<uint8_t> prove_all(ola_ctx* ctx, const std::vector<int>& table_ids, const std::vector<const uint64_t*>& traces, bool on_device,
                               const std::vector<uint32_t>& log_ns, const std::vector<uint64_t>& compress_challenges, const Config& cfg,
                               TranscriptHost* transcript_host, const LateTable* late) {
    System sys = make_system(table_ids);
    OLA_CHECK(!late || (on_device && late->index < table_ids.size()), OLA_ERR_INVALID_ARG, "a late table is a device-resident table of the system");
    OLA_CHECK(compress_challenges.empty() || compress_challenges.size() == sys.tables.size(), OLA_ERR_INVALID_ARG, "one compress challenge per table");
    for (size_t i = 0; i < compress_challenges.size(); ++i)
        if (sys.tables[i].id == T_BITWISE || sys.tables[i].id == T_PROGRAM)
            sys.tables[i].compress_challenge = sys.compress_challenges[i] = gl::canon(compress_challenges[i]);
    const size_t T = sys.tables.size();
    OLA_CHECK(traces.size() == T && log_ns.size() == T, OLA_ERR_INVALID_ARG, "one trace per table");
    std::vector<std::unique_ptr<DevBuf>> d_vals(T);
    std::vector<std::unique_ptr<BatchHolder>> commits(T);
    Challenger ch(ctx->hasher, transcript_host);
    // host traces: every table's upload is queued on the copy stream up front, so table i+1 crosses PCIe while table i is
    // being committed (LDE + Poseidon) on the context stream.  With several ranks each uploads 1/world of the columns
    // over its own PCIe link (into `slices`) and one all-gather over NVLink replicates the table.
    const bool prefetch = !on_device;
    const bool sharded_upload = prefetch && ctx->world > 1;
    const bool side_comm = sharded_upload && nccl::has_side_comm(ctx);
    std::vector<std::unique_ptr<DevBuf>> slices(T);
    // The 12 trace commitments are independent (their caps are observed afterwards, in table order), so they are
    // uploaded and committed smallest first: the first commit starts after a short copy and the big tables cross
    // PCIe behind the hashing of the small ones instead of in front of everything (the CPU table alone is 57 ms of copy).
    std::vector<size_t> order(T);
    for (size_t i = 0; i < T; ++i) order[i] = i;
    if (prefetch)
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) {
            return ((size_t)sys.tables[x].columns << log_ns[x]) < ((size_t)sys.tables[y].columns << log_ns[y]);
        });
    if (late) {  // committed last: whatever completes it runs beside the other commitments
        order.erase(std::find(order.begin(), order.end(), late->index));
        order.push_back(late->index);
    }
    std::vector<cudaEvent_t> uploaded(T, nullptr);
    struct EventGuard {
        std::vector<cudaEvent_t>& v;
        ~EventGuard() {
            for (auto e : v)
                if (e) cudaEventDestroy(e);
        }
    } event_guard{uploaded};
    try {  // from the first queued copy on: no copy may outlive the buffers released by unwinding
    if (prefetch) {
        ensure_copy_stream(ctx);
        cudaEvent_t allocated;
        OLA_CUDA(cudaEventCreateWithFlags(&allocated, cudaEventDisableTiming));
        for (size_t i = 0; i < T; ++i) {
            OLA_CHECK(log_ns[i] + Config::rate_bits <= 32, OLA_ERR_INVALID_ARG, "trace too long for the field's two-adicity");
            const size_t n = (size_t)1 << log_ns[i], cols = (size_t)sys.tables[i].columns;
            if (sharded_upload) {
                const size_t per = (cols + ctx->world - 1) / ctx->world;  // padded to equal contributions
                d_vals[i].reset(new DevBuf(per * ctx->world * n));
                slices[i].reset(new DevBuf(per * n));
                const size_t lo = std::min(cols, (size_t)ctx->rank * per), hi = std::min(cols, lo + per);
                if (hi - lo < per) OLA_CUDA(cudaMemsetAsync(slices[i]->p, 0, per * n * 8, ctx->stream));
            } else {
                d_vals[i].reset(new DevBuf(n * cols));
            }
        }
        cudaError_t e = cudaEventRecord(allocated, ctx->stream);  // the stream-ordered allocations above precede the copies
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, allocated, 0);
        cudaEventDestroy(allocated);
        OLA_CUDA(e);
        for (size_t oi = 0; oi < T; ++oi) {
            const size_t i = order[oi];
            const size_t n = (size_t)1 << log_ns[i], cols = (size_t)sys.tables[i].columns;
            if (sharded_upload) {
                const size_t per = (cols + ctx->world - 1) / ctx->world;
                const size_t lo = std::min(cols, (size_t)ctx->rank * per), hi = std::min(cols, lo + per);
                if (hi > lo) OLA_CUDA(cudaMemcpyAsync(slices[i]->p, traces[i] + lo * n, (hi - lo) * n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
                if (side_comm) {
                    // canonicalise and replicate the table right behind its upload, on the copy stream and the second NCCL
                    // communicator: these all-gathers overlap the commitments of the tables in front
                    const unsigned blocks = (unsigned)std::min<size_t>((per * n + 255) / 256, 148 * 16);
                    canon_copy_kernel<<<blocks, 256, 0, ctx->copy_stream>>>(slices[i]->p, slices[i]->p, per * n);
                    check_launch("canon_copy_kernel");
                    count_launch(ctx);
                    nccl::allgather_side(ctx, slices[i]->p, d_vals[i]->p, per * n * 8, ctx->copy_stream);
                }
            } else {
                OLA_CUDA(cudaMemcpyAsync(d_vals[i]->p, traces[i], n * cols * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            }
            OLA_CUDA(cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming));
            OLA_CUDA(cudaEventRecord(uploaded[i], ctx->copy_stream));
        }
    }
    {
        for (size_t oi = 0; oi < T; ++oi) {
            const size_t i = order[oi];
            const size_t n = (size_t)1 << log_ns[i], cnt = n * sys.tables[i].columns;
            OLA_CHECK(log_ns[i] + Config::rate_bits <= 32, OLA_ERR_INVALID_ARG, "trace too long for the field's two-adicity");
            if (sharded_upload && side_comm) {
                OLA_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded[i], 0));  // uploaded, canonicalised and all-gathered
            } else if (sharded_upload) {
                const size_t per = ((size_t)sys.tables[i].columns + ctx->world - 1) / ctx->world;
                OLA_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded[i], 0));
                canon_copy(ctx, slices[i]->p, slices[i]->p, per * n);
                comm_allgather(ctx, slices[i]->p, d_vals[i]->p, per * n * 8);
            } else if (prefetch) {
                OLA_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded[i], 0));
                canon_copy(ctx, d_vals[i]->p, d_vals[i]->p, cnt);  // the Z kernels read these values with canonical-input arithmetic
            } else {
                if (late && late->index == i) sys.tables[i].compress_challenge = sys.compress_challenges[i] = gl::canon(late->finish());
                d_vals[i].reset(new DevBuf(cnt));
                OLA_CUDA(cudaMemcpyAsync(d_vals[i]->p, traces[i], cnt * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                canon_copy(ctx, d_vals[i]->p, d_vals[i]->p, cnt);
            }
            commits[i].reset(new BatchHolder());
            commits[i]->b = commit(ctx, d_vals[i]->p, sys.tables[i].columns, log_ns[i], false);
        }
    }
    } catch (...) {
        if (prefetch) cudaStreamSynchronize(ctx->copy_stream);  // no copy may outlive the buffers released by unwinding
        throw;
    }
    if (sharded_upload) {
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));  // the all-gathers have consumed the slices
        if (side_comm) OLA_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        slices.clear();
    }
    ch.at(STAGE_TRACE_CAPS);
    for (size_t i = 0; i < T; ++i) ch.observe_cap(batch_cap(ctx, commits[i]->b));
    ch.at(STAGE_CTL_CHALLENGES);
    // cross_table_lookup_data: challenges, then Z instances per table in registry order
    std::vector<Challenge> ctl_ch;
    for (uint32_t k = 0; k < Config::num_challenges; ++k) {
        F b = ch.get_challenge();
        F g = ch.get_challenge();
        ctl_ch.push_back({b, g});
    }
    std::vector<std::vector<CtlInstance>> per_table(T);
    for (auto& ctl : sys.ctls)
        for (auto& c : ctl_ch) {
            for (auto& lt : ctl.looking) per_table[lt.table].push_back({c, &lt});
            if (ctl.has_looked) per_table[ctl.looked.table].push_back({c, &ctl.looked});
        }
    Writer w;
    w.raw_hashes = ctx->hasher == OLA_HASH_BLAKE3;
    w.u32((uint32_t)T);
    for (size_t i = 0; i < T; ++i) {
        ch.table = sys.tables[i].id;
        StarkProof p = prove_single_table(ctx, sys.tables[i], cfg, d_vals[i]->p, log_ns[i], commits[i]->b, per_table[i], ch);
        w.proof(p);
        commits[i].reset();  // tables are proven sequentially: free this table's HBM before the next
        d_vals[i].reset();
    }
    w.field_vec(sys.compress_challenges);
    return w.buf;
}

}  // namespace stark
}  // namespace ola
