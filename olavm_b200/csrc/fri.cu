// FRI on the device: batch composition, quotient by (X - z), extension-field LDE, per-layer commit and fold,
// proof-of-work grind and query openings.
//
// Replaces:
//   plonky2/plonky2/src/fri/oracle.rs:167-241   PolynomialBatch::prove_openings
//   plonky2/plonky2/src/fri/prover.rs:20-204    fri_proof, fri_committed_trees, fri_proof_of_work, query rounds
//   plonky2/field/src/polynomial/division.rs:74-87  divide_by_linear
//   plonky2/plonky2/src/plonk/plonk_common.rs:116-128  reduce_with_powers
//
// Extension-field polynomials live PLANAR in HBM ([2][len]: all c0, then all c1): the 2^k-th roots of unity and
// the coset shifts are base-field elements (goldilocks_extensions.rs:27), so an extension NTT is two base NTTs
// and reuses the batched kernels of ntt.cu unchanged.  The transcript stays on the host (stark_types.h);
// only caps, the final polynomial, the PoW nonce and the opened rows ever cross PCIe.
#include "fri.h"

#include "air/air_common.cuh"
#include "hasher.h"

#include "batch.h"
#include "ntt.h"

namespace ola {
namespace fri {

using stark::Config;
using stark::E;
using stark::F;

static constexpr int MAX_POLYS = 1024;

// ---- composition: F_b[j] = sum_i alpha^i * f_{b,i}[j] for the three opening batches at once ----------------
struct ComposeDesc {
    const uint64_t* ptr[MAX_POLYS];  // coefficient column of polynomial p
    int16_t idx[3][MAX_POLYS];       // position of p inside batch b, or -1
};
// Each batch's sum is accumulated UNREDUCED (air::Wide: 6 instructions per multiply-accumulate instead of the 28 of a reduced
// multiply + add) and reduced once per output: the kernel reads every committed coefficient once and was bound by the
// integer pipes, not by HBM (10.6 ms for 10.4 GB before).
__global__ void __launch_bounds__(128) compose_kernel(const ComposeDesc* __restrict__ d, int npolys, size_t n, const uint64_t* __restrict__ apow /* [2][maxlen] */,
                                                      size_t apow_stride, uint64_t* __restrict__ comp /* [3][2][n] */) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    air::Wide acc[3][2];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        acc[b][0].clear();
        acc[b][1].clear();
    }
    // four polynomials per trip: the pointer fetch and the coefficient load of a polynomial are dependent loads, so the
    // loads of a group are issued together before any of them is consumed (memory-level parallelism; the kernel streams
    // every committed coefficient once)
    int p = 0;
    for (; p + 4 <= npolys; p += 4) {
        const uint64_t* q0 = d->ptr[p];
        const uint64_t* q1 = d->ptr[p + 1];
        const uint64_t* q2 = d->ptr[p + 2];
        const uint64_t* q3 = d->ptr[p + 3];
        const uint64_t v[4] = {__ldg(q0 + j), __ldg(q1 + j), __ldg(q2 + j), __ldg(q3 + j)};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int k = d->idx[b][p + u];
                if (k >= 0) {
                    acc[b][0].mac(v[u], __ldg(apow + k));
                    acc[b][1].mac(v[u], __ldg(apow + apow_stride + k));
                }
            }
        }
    }
    for (; p < npolys; ++p) {
        const uint64_t v = __ldg(d->ptr[p] + j);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int k = d->idx[b][p];
            if (k >= 0) {
                acc[b][0].mac(v, __ldg(apow + k));
                acc[b][1].mac(v, __ldg(apow + apow_stride + k));
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        comp[((size_t)b * 2) * n + j] = acc[b][0].reduce();
        comp[((size_t)b * 2 + 1) * n + j] = acc[b][1].reduce();
    }
}

// ---- quotient by linear: S_b[m] = sum_{k >= m} c_k z_b^(k-m);  final[m] = sum_b w_b S_b[m] (m >= 1), final[0] = 0
static constexpr int QL_RUN = 8;
static constexpr int QL_THREADS = 256;
static constexpr int QL_CHUNK = QL_RUN * QL_THREADS;
struct QlParams {
    gl::ext2 z[3], w[3];
    gl::ext2 zr[3][9];   // z^(RUN * 2^k), k = 0..8
    gl::ext2 zchunk[3];  // z^CHUNK
    int nbatches;
};

__device__ __forceinline__ gl::ext2 ld2(const uint64_t* base, size_t n, size_t j) { return gl::make2(base[j], base[n + j]); }

// phase 1: chunk value C_s = sum_{k in chunk s} c_k z^(k - s*CHUNK)
__global__ void __launch_bounds__(QL_THREADS) ql_chunk_kernel(const uint64_t* __restrict__ comp, size_t n, QlParams p, uint64_t* __restrict__ chunkv /* [3][2][nchunks] */,
                                                           size_t nchunks) {
    __shared__ gl::ext2 sh[QL_THREADS];
    const int b = blockIdx.y, t = threadIdx.x;
    const uint64_t* c = comp + (size_t)b * 2 * n;
    size_t start = (size_t)blockIdx.x * QL_CHUNK + (size_t)t * QL_RUN;
    gl::ext2 acc = gl::make2(0, 0);
    for (int k = QL_RUN - 1; k >= 0; --k) {
        size_t j = start + k;
        gl::ext2 v = j < n ? ld2(c, n, j) : gl::make2(0, 0);
        acc = gl::add(gl::mul(acc, p.z[b]), v);
    }
    sh[t] = acc;
    __syncthreads();
    for (int k = 0; (1 << k) < QL_THREADS; ++k) {
        gl::ext2 v = sh[t];
        if ((t & ((2 << k) - 1)) == 0) v = gl::add(v, gl::mul(p.zr[b][k], sh[t + (1 << k)]));
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    if (t == 0) {
        chunkv[((size_t)b * 2) * nchunks + blockIdx.x] = sh[0].c0;
        chunkv[((size_t)b * 2 + 1) * nchunks + blockIdx.x] = sh[0].c1;
    }
}
// phase 2: carry T_s = S[(s+1)*CHUNK] = C_{s+1} + z^CHUNK * T_{s+1}; one thread per batch
// carries between chunks: T_s = sum_{u > s} C_u zc^(u - s - 1) (zc = z^CHUNK), written over C_s.  One CTA per batch: every
// thread owns a run of consecutive chunks, the runs are combined by a suffix scan over the affine maps
// carry -> A + M carry (A = the run's own Horner value, M = zc^run length), then each thread replays its run with the carry
// that enters it.  (A single thread per batch walking all chunks took 0.7 ms per 2^22-row table: pure latency.)
static constexpr int QLC_THREADS = 1024;
__global__ void __launch_bounds__(QLC_THREADS) ql_carry_kernel(uint64_t* chunkv, size_t nchunks, QlParams p) {
    __shared__ gl::ext2 sm[QLC_THREADS], sa[QLC_THREADS];
    const int b = blockIdx.x, t = threadIdx.x;
    uint64_t* c0 = chunkv + ((size_t)b * 2) * nchunks;
    uint64_t* c1 = c0 + nchunks;
    const gl::ext2 zc = p.zchunk[b];
    const size_t per = (nchunks + QLC_THREADS - 1) / QLC_THREADS;
    const size_t lo = (size_t)t * per < nchunks ? (size_t)t * per : nchunks, hi = lo + per < nchunks ? lo + per : nchunks;
    gl::ext2 A = gl::make2(0, 0), M = gl::make2(1, 0);
    for (size_t s = hi; s-- > lo;) {
        A = gl::add(gl::make2(c0[s], c1[s]), gl::mul(zc, A));
        M = gl::mul(M, zc);
    }
    sm[t] = M;
    sa[t] = A;
    __syncthreads();
    // inclusive suffix scan: (M, A)[t] <- composition of the runs t, t + 1, ... (the lower run is applied last)
    for (int d = 1; d < QLC_THREADS; d <<= 1) {
        gl::ext2 m = sm[t], a = sa[t];
        if (t + d < QLC_THREADS) {
            a = gl::add(a, gl::mul(m, sa[t + d]));
            m = gl::mul(m, sm[t + d]);
        }
        __syncthreads();
        sm[t] = m;
        sa[t] = a;
        __syncthreads();
    }
    gl::ext2 carry = t + 1 < QLC_THREADS ? sa[t + 1] : gl::make2(0, 0);  // what the runs above hand down (they start from 0)
    for (size_t s = hi; s-- > lo;) {
        const gl::ext2 cs = gl::make2(c0[s], c1[s]);
        c0[s] = carry.c0;  // overwrite C_s with T_s
        c1[s] = carry.c1;
        carry = gl::add(cs, gl::mul(zc, carry));
    }
}
__global__ void __launch_bounds__(QL_THREADS) ql_emit_kernel(const uint64_t* __restrict__ comp, size_t n, QlParams p, const uint64_t* __restrict__ carry,
                                                          size_t nchunks, uint64_t* __restrict__ final_poly /* [2][n] */) {
    __shared__ gl::ext2 sh[QL_THREADS + 1];
    const int t = threadIdx.x;
    size_t start = (size_t)blockIdx.x * QL_CHUNK + (size_t)t * QL_RUN;
    gl::ext2 out[QL_RUN];
#pragma unroll
    for (int k = 0; k < QL_RUN; ++k) out[k] = gl::make2(0, 0);
    for (int b = 0; b < p.nbatches; ++b) {
        const uint64_t* c = comp + (size_t)b * 2 * n;
        gl::ext2 cv[QL_RUN];
        gl::ext2 acc = gl::make2(0, 0);
        for (int k = QL_RUN - 1; k >= 0; --k) {
            size_t j = start + k;
            cv[k] = j < n ? ld2(c, n, j) : gl::make2(0, 0);
            acc = gl::add(gl::mul(acc, p.z[b]), cv[k]);
        }
        __syncthreads();
        sh[t] = acc;
        if (t == 0) sh[QL_THREADS] = gl::make2(carry[((size_t)b * 2) * nchunks + blockIdx.x], carry[((size_t)b * 2 + 1) * nchunks + blockIdx.x]);
        __syncthreads();
        // inclusive suffix scan over the QL_THREADS + 1 entries: A_t = sum_{t' >= t} R_t' z^(RUN (t' - t))
        for (int k = 0; k < 9; ++k) {
            gl::ext2 v = sh[t];
            int o = t + (1 << k);
            if (o <= QL_THREADS) v = gl::add(v, gl::mul(p.zr[b][k], sh[o]));
            gl::ext2 vl = sh[QL_THREADS];
            __syncthreads();
            sh[t] = v;
            if (t == 0) sh[QL_THREADS] = vl;
            __syncthreads();
        }
        gl::ext2 s = sh[t + 1];  // S at the end of this thread's run
        for (int k = QL_RUN - 1; k >= 0; --k) {
            s = gl::add(gl::mul(s, p.z[b]), cv[k]);  // S[start + k]
            out[k] = gl::add(out[k], gl::mul(p.w[b], s));
        }
    }
#pragma unroll
    for (int k = 0; k < QL_RUN; ++k) {
        size_t j = start + k;
        if (j < n) {
            gl::ext2 v = j == 0 ? gl::make2(0, 0) : out[k];
            final_poly[j] = v.c0;
            final_poly[n + j] = v.c1;
        }
    }
}

// ---- FRI layer leaves: leaf i = flatten(values[16 i .. 16 i + 16)) -> digest (hash_no_pad of 32 elements) ----
// (leaves [leaf_first, leaf_first + leaf_count) of the layer; digests[4 i] is the digest of leaf leaf_first + i: a rank of
// the leaf-range-sharded commit phase hashes its own range)
__global__ void __launch_bounds__(128, 5) fri_leaves_kernel(const uint64_t* __restrict__ vals /* [2][len] */, size_t len, int arity, size_t leaf_first,
                                                            size_t leaf_count, uint64_t* __restrict__ digests) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= leaf_count) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k] = 0;
    const uint64_t* c0 = vals + (leaf_first + i) * arity;
    const uint64_t* c1 = vals + len + (leaf_first + i) * arity;
    for (int e = 0; e < arity; e += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (e + k < arity) {
                s[2 * k] = c0[e + k];
                s[2 * k + 1] = c1[e + k];
            }
        }
        poseidon::permute(s);
    }
    ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + 4 * i);
    d[0] = make_ulonglong2(s[0], s[1]);
    d[1] = make_ulonglong2(s[2], s[3]);
}

// ---- fold: out[j] = sum_{i < arity} beta^i c[arity*j + i] ----
__global__ void fold_kernel(const uint64_t* __restrict__ in /* [2][m] */, size_t m, int arity, gl::ext2 beta, uint64_t* __restrict__ out /* [2][m/arity] */) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t mo = m / arity;
    if (j >= mo) return;
    gl::ext2 s = gl::make2(0, 0);
    for (int i = arity - 1; i >= 0; --i) s = gl::add(gl::mul(s, beta), gl::make2(in[j * arity + i], in[m + j * arity + i]));
    out[j] = s.c0;
    out[mo + j] = s.c1;
}

// ---- proof of work: smallest nonce with hash_no_pad(h0..h3, nonce)[0] < 2^(64 - pow_bits) (fri/prover.rs:126-148) ----
__global__ void __launch_bounds__(128, 5) pow_kernel(uint64_t h0, uint64_t h1, uint64_t h2, uint64_t h3, uint64_t base, unsigned long long* best) {
    uint64_t nonce = base + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // CTAs are dispatched in index order: once an earlier one has recorded a valid nonce, the rest of the window (the
    // expected first hit is ~2^16 of its 2^20 nonces) has nothing smaller to offer and retires without hashing
    if (*reinterpret_cast<volatile unsigned long long*>(best) < (unsigned long long)nonce) return;
    uint64_t s[12] = {h0, h1, h2, h3, nonce, 0, 0, 0, 0, 0, 0, 0};
    poseidon::permute(s);
    if ((s[0] >> (64 - Config::pow_bits)) == 0) atomicMin(best, (unsigned long long)nonce);
}

// ---- query gathers ----
// rows of a column-major matrix at the query indices: out[q][c] = m[c*stride + idx[q]].  The matrix holds leaves
// [leaf_lo, leaf_lo + stride) of the tree (a coset shard; the whole tree when leaf_lo = 0 and stride = nleaves): queries
// this rank does not own produce zeros, to be summed with the owner's answer.
__global__ void gather_query_rows_kernel(const uint64_t* __restrict__ m, size_t stride, size_t ncols, const uint32_t* __restrict__ idx, int shift, int nq,
                                         size_t leaf_lo, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nq * ncols) return;
    size_t q = i / ncols, c = i % ncols;
    size_t leaf = idx[q] >> shift;
    out[i] = (leaf >= leaf_lo && leaf - leaf_lo < stride) ? m[c * stride + (leaf - leaf_lo)] : 0;
}
// FRI layer leaf (arity ext values interleaved) at idx[q] >> shift: out[q][2*e + comp]
// (vals = the layer values of leaves [leaf_lo, leaf_lo + nleaves_local), planes `len` apart; leaves of other ranks -> 0)
__global__ void gather_query_ext_kernel(const uint64_t* __restrict__ vals, size_t len, int arity, const uint32_t* __restrict__ idx, int shift, int nq,
                                        size_t leaf_lo, size_t nleaves_local, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nq * arity * 2) return;
    size_t q = i / (2 * arity), r = i % (2 * arity);
    size_t e = r / 2, comp = r % 2;
    size_t leaf = idx[q] >> shift;
    if (leaf < leaf_lo || leaf - leaf_lo >= nleaves_local) {
        out[i] = 0;
        return;
    }
    out[i] = vals[comp * len + (leaf - leaf_lo) * arity + e];
}
// Merkle paths: out[q][j][w] = nodes[(((nleaves + leaf) >> j) ^ 1) * 4 + w], leaf = idx[q] >> shift
// (nodes = the heap-ordered subtree over leaves [leaf_lo, leaf_lo + nleaves); paths stop at the cap, inside the subtree)
__global__ void gather_query_paths_kernel(const uint64_t* __restrict__ nodes, size_t nleaves, const uint32_t* __restrict__ idx, int shift, int nq, int nsib,
                                          size_t leaf_lo, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nq * nsib * 4) return;
    size_t q = i / ((size_t)nsib * 4), r = i % ((size_t)nsib * 4);
    size_t j = r / 4, w = r % 4;
    size_t leaf = idx[q] >> shift;
    if (leaf < leaf_lo || leaf - leaf_lo >= nleaves) {
        out[i] = 0;
        return;
    }
    leaf -= leaf_lo;
    out[i] = nodes[((((nleaves + leaf) >> j) ^ 1) << 2) + w];
}

// ---------------------------------------------------------------------------------------------------------
struct DevMem {  // RAII cudaMalloc
    uint64_t* p = nullptr;
    DevMem() {}
    explicit DevMem(size_t n) { dev_alloc(&p, n); }
    void alloc(size_t n) {
        release();
        dev_alloc(&p, n);
    }
    void release() {
        if (p) ola::dev_free(p);
        p = nullptr;
    }
    ~DevMem() { release(); }
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
};

static gl::ext2 epow(gl::ext2 b, uint64_t e) { return gl::pow(b, e); }

stark::FriProof prove_openings(ola_ctx* ctx, const Instance& inst, const std::vector<const ola_batch*>& oracles, stark::Challenger& ch,
                               uint32_t degree_bits) {
    const size_t n = (size_t)1 << degree_bits;
    const uint32_t lde_bits = degree_bits + Config::rate_bits;
    const size_t L = (size_t)1 << lde_bits;
    const std::vector<uint32_t> arities = stark::fri_arities(degree_bits);
    OLA_CHECK(inst.batches.size() <= 3 && !inst.batches.empty(), OLA_ERR_INTERNAL, "FRI instance: 1..3 opening batches supported");
    cudaStream_t st = ctx->stream;

    ch.at(stark::STAGE_FRI_ALPHA);
    E alpha = ch.get_ext();

    // ---- polynomial list and per-batch alpha-power indices
    std::vector<ComposeDesc> hd(1);
    memset(hd[0].idx, 0xff, sizeof(hd[0].idx));
    std::vector<size_t> base(oracles.size() + 1, 0);
    for (size_t o = 0; o < oracles.size(); ++o) base[o + 1] = base[o] + oracles[o]->ncols;
    const int npolys = (int)base.back();
    OLA_CHECK(npolys <= MAX_POLYS, OLA_ERR_INTERNAL, "too many committed polynomials for one FRI instance");
    for (size_t o = 0; o < oracles.size(); ++o)
        for (size_t c = 0; c < oracles[o]->ncols; ++c) hd[0].ptr[base[o] + c] = oracles[o]->d_coeffs + c * n;
    size_t maxlen = 1;
    for (size_t b = 0; b < inst.batches.size(); ++b) {
        auto& polys = inst.batches[b].polys;
        maxlen = std::max(maxlen, polys.size());
        for (size_t i = 0; i < polys.size(); ++i) hd[0].idx[b][base[polys[i].first] + polys[i].second] = (int16_t)i;
    }
    std::vector<uint64_t> hap(2 * maxlen);
    {
        E a = gl::make2(1, 0);
        for (size_t i = 0; i < maxlen; ++i) {
            hap[i] = a.c0;
            hap[maxlen + i] = a.c1;
            a = gl::mul(a, alpha);
        }
    }
    DevMem d_desc((sizeof(ComposeDesc) + 7) / 8), d_apow(2 * maxlen), d_comp(3 * 2 * n);
    OLA_CUDA(cudaMemcpyAsync(d_desc.p, hd.data(), sizeof(ComposeDesc), cudaMemcpyHostToDevice, st));
    OLA_CUDA(cudaMemcpyAsync(d_apow.p, hap.data(), hap.size() * 8, cudaMemcpyHostToDevice, st));
    {
        Launch lz(ctx, "fri_compose");
        compose_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const ComposeDesc*)d_desc.p, npolys, n, d_apow.p, maxlen, d_comp.p);
    }
    check_launch("compose_kernel");

    // ---- quotients by (X - z_b), weights w_b = alpha^(sum of later batch lengths)  (oracle.rs:193-214)
    QlParams qp;
    memset(&qp, 0, sizeof(qp));
    qp.nbatches = (int)inst.batches.size();
    for (int b = 0; b < qp.nbatches; ++b) {
        qp.z[b] = inst.batches[b].point;
        size_t later = 0;
        for (int b2 = b + 1; b2 < qp.nbatches; ++b2) later += inst.batches[b2].polys.size();
        qp.w[b] = epow(alpha, later);
        E zr = epow(qp.z[b], QL_RUN);
        for (int k = 0; k < 9; ++k) {
            qp.zr[b][k] = zr;
            zr = gl::mul(zr, zr);
        }
        qp.zchunk[b] = epow(qp.z[b], QL_CHUNK);
    }
    const size_t nchunks = (n + QL_CHUNK - 1) / QL_CHUNK;
    DevMem d_chunk(3 * 2 * nchunks), d_final(2 * n);
    {
        Launch lz(ctx, "fri_divide_chunks");
        ql_chunk_kernel<<<dim3((unsigned)nchunks, (unsigned)qp.nbatches), QL_THREADS, 0, st>>>(d_comp.p, n, qp, d_chunk.p, nchunks);
    }
    check_launch("ql_chunk_kernel");
    {
        Launch lz(ctx, "fri_divide_carry");
        ql_carry_kernel<<<(unsigned)qp.nbatches, QLC_THREADS, 0, st>>>(d_chunk.p, nchunks, qp);
    }
    check_launch("ql_carry_kernel");
    {
        Launch lz(ctx, "fri_divide_emit");
        ql_emit_kernel<<<(unsigned)nchunks, QL_THREADS, 0, st>>>(d_comp.p, n, qp, d_chunk.p, nchunks, d_final.p);
    }
    check_launch("ql_emit_kernel");
    d_comp.release();

    // ---- commit phase (fri_committed_trees, prover.rs:72-121)
    struct Layer {
        DevMem vals, nodes;  // nodes: heap-ordered tree over leaves [leaf_lo, leaf_lo + nloc)
        size_t len = 0, len_loc = 0, nloc = 0, leaf_lo = 0;  // vals: [2][len_loc], the values of this rank's leaves
        bool sharded = false;
    };
    std::vector<std::unique_ptr<Layer>> layers;
    stark::FriProof proof;
    DevMem coeffs;  // current coefficients [2][m]
    coeffs.p = d_final.p;
    d_final.p = nullptr;
    size_t m = n;  // number of (possibly) non-zero coefficients
    F shift = gl::GEN;
    for (size_t li = 0; li < arities.size(); ++li) {
        const int arity = 1 << arities[li];
        std::unique_ptr<Layer> ly(new Layer());
        const uint32_t mbits = (uint32_t)__builtin_ctzll(m);
        ly->len = m << Config::rate_bits;
        const size_t nleaves = ly->len / arity;
        const size_t ncap = (size_t)1 << Config::cap_height;
        OLA_CHECK(nleaves >= ncap, OLA_ERR_INTERNAL, "FRI layer smaller than the Merkle cap");
        // Multi-GPU: a large layer is sharded by leaf range = LDE cosets: rank r evaluates cosets [r 8 / world, ...) of the
        // layer polynomial, hashes those leaves and reduces the ncap / world cap subtrees above them; one all-gather
        // assembles the cap, and query openings are answered by the owner of the leaf.  Small layers are replicated.
        // (OLA_FRI_SHARD_MIN_LEAVES lowers the threshold so that tests reach the sharded path with small tables)
        const char* env_min = getenv("OLA_FRI_SHARD_MIN_LEAVES");
        const size_t shard_min = env_min ? (size_t)atoll(env_min) : ((size_t)1 << 14);
        const int ncosets = 1 << Config::rate_bits;
        ly->sharded = ctx->world > 1 && nleaves >= shard_min && nleaves >= ncap * (size_t)ctx->world && ncap % (size_t)ctx->world == 0 &&
                      ncosets % ctx->world == 0 && (m % (size_t)arity) == 0;
        ly->nloc = ly->sharded ? nleaves / (size_t)ctx->world : nleaves;
        ly->leaf_lo = ly->sharded ? (size_t)ctx->rank * ly->nloc : 0;
        ly->len_loc = ly->sharded ? ly->len / (size_t)ctx->world : ly->len;
        const size_t ncap_loc = ly->sharded ? ncap / (size_t)ctx->world : ncap;
        ly->vals.alloc(2 * ly->len_loc);
        {
            ntt::FwdDesc d;
            d.src = coeffs.p;
            d.src_col_stride = m;
            d.dst = ly->vals.p;
            d.dst_col_stride = ly->len_loc;
            d.dst_coset_stride = m;
            d.ncols = 2;
            d.log_n = (int)mbits;
            d.coset_bits = (int)Config::rate_bits;
            if (ly->sharded) {
                d.coset_count = ncosets / ctx->world;
                d.coset_first = ctx->rank * d.coset_count;
            }
            d.shift = shift;
            d.tag_strided = "fri_lde_strided";
            d.tag_contig = "fri_lde_contig";
            ntt::forward(ctx, d);
        }
        ly->nodes.alloc(2 * ly->nloc * 4);
        if (ctx->hasher == OLA_HASH_BLAKE3) {
            blake3::fri_leaves(ctx, ly->vals.p, ly->len_loc, arity, 0, ly->nloc, ly->nodes.p + 4 * ly->nloc);
        } else {
            Launch lz(ctx, "fri_leaves");
            fri_leaves_kernel<<<(unsigned)((ly->nloc + 127) / 128), 128, 0, st>>>(ly->vals.p, ly->len_loc, arity, 0, ly->nloc, ly->nodes.p + 4 * ly->nloc);
        }
        check_launch("fri_leaves_kernel");
        hasher::merkle_levels(ctx, ly->nodes.p, ly->nloc, ncap_loc);
        stark::Cap cap(ncap);
        if (ly->sharded) {
            DevMem d_cap(ncap * 4);
            comm_allgather(ctx, ly->nodes.p + 4 * ncap_loc, d_cap.p, ncap_loc * 32);
            OLA_CUDA(cudaMemcpyAsync(cap.data(), d_cap.p, ncap * 32, cudaMemcpyDeviceToHost, st));
            OLA_CUDA(cudaStreamSynchronize(st));
        } else {
            OLA_CUDA(cudaMemcpyAsync(cap.data(), ly->nodes.p + 4 * ncap, ncap * 32, cudaMemcpyDeviceToHost, st));
            OLA_CUDA(cudaStreamSynchronize(st));
        }
        ch.at(stark::STAGE_FRI_LAYER_CAP);
        ch.observe_cap(cap);
        proof.commit_caps.push_back(cap);
        ch.at(stark::STAGE_FRI_BETA);
        E beta = ch.get_ext();
        DevMem folded(2 * (m / arity));
        {
            Launch lz(ctx, "fri_fold");
            fold_kernel<<<(unsigned)((m / arity + 127) / 128), 128, 0, st>>>(coeffs.p, m, arity, beta, folded.p);
        }
        check_launch("fold_kernel");
        OLA_CUDA(cudaStreamSynchronize(st));
        coeffs.release();
        coeffs.p = folded.p;
        folded.p = nullptr;
        m /= arity;
        shift = gl::pow(shift, (uint64_t)arity);
        layers.push_back(std::move(ly));
    }
    // final polynomial (coeffs.truncate(len >> rate_bits): exactly the m coefficients kept here)
    {
        std::vector<uint64_t> h(2 * m);
        OLA_CUDA(cudaMemcpyAsync(h.data(), coeffs.p, 2 * m * 8, cudaMemcpyDeviceToHost, st));
        OLA_CUDA(cudaStreamSynchronize(st));
        ch.at(stark::STAGE_FRI_FINAL_POLY);
        for (size_t i = 0; i < m; ++i) {
            proof.final_poly.push_back(gl::make2(h[i], h[m + i]));
            ch.observe_ext(proof.final_poly.back());
        }
    }
    // ---- proof of work
    {
        ch.at(stark::STAGE_FRI_POW);
        stark::Hash h = ch.get_hash();
        DevMem d_best(1);
        unsigned long long init = ~0ULL;
        unsigned long long best = init;
        const uint64_t window = 1 << 20;
        for (uint64_t basev = 0; best == init; basev += window) {
            OLA_CUDA(cudaMemcpyAsync(d_best.p, &init, 8, cudaMemcpyHostToDevice, st));
            {
                Launch lz(ctx, "fri_pow");
                pow_kernel<<<(unsigned)(window / 128), 128, 0, st>>>(h.e[0], h.e[1], h.e[2], h.e[3], basev, (unsigned long long*)d_best.p);
            }
            check_launch("pow_kernel");
            OLA_CUDA(cudaMemcpyAsync(&best, d_best.p, 8, cudaMemcpyDeviceToHost, st));
            OLA_CUDA(cudaStreamSynchronize(st));
            OLA_CHECK(basev < ((uint64_t)1 << 40), OLA_ERR_INTERNAL, "Proof of work failed. This is highly unlikely!");
        }
        proof.pow_witness = (F)best;
    }
    // ---- query phase (fri_prover_query_rounds, prover.rs:150-204): one batched gather per tree
    const int nq = (int)Config::num_queries;
    std::vector<uint32_t> idx(nq);
    ch.at(stark::STAGE_FRI_QUERY_INDICES);
    {
        const std::vector<stark::F> qs = ch.get_challenges((size_t)nq);
        for (int q = 0; q < nq; ++q) idx[q] = (uint32_t)(qs[q] % L);
    }
    DevMem d_idx((nq + 1) / 2 + 1);
    OLA_CUDA(cudaMemcpyAsync(d_idx.p, idx.data(), nq * 4, cudaMemcpyHostToDevice, st));
    // output layout: per oracle [nq][ncols] rows then [nq][nsib][4] paths; per layer [nq][2*arity] then paths
    size_t total = 0;
    const int nsib0 = (int)lde_bits - (int)Config::cap_height;
    std::vector<size_t> off_rows(oracles.size()), off_paths(oracles.size());
    for (size_t o = 0; o < oracles.size(); ++o) {
        off_rows[o] = total;
        total += (size_t)nq * oracles[o]->ncols;
        off_paths[o] = total;
        total += (size_t)nq * nsib0 * 4;
    }
    // the layers' rows and Merkle paths follow: owner-answered like the oracle openings when the layer is sharded
    std::vector<size_t> off_lrows(layers.size()), off_lpaths(layers.size());
    std::vector<int> lshift(layers.size()), lnsib(layers.size());
    {
        int sh = 0;
        uint32_t bits = lde_bits;
        for (size_t li = 0; li < layers.size(); ++li) {
            sh += (int)arities[li];
            bits -= arities[li];
            lshift[li] = sh;
            lnsib[li] = (int)bits - (int)Config::cap_height;
            off_lrows[li] = total;
            total += (size_t)nq * 2 * (1 << arities[li]);
            off_lpaths[li] = total;
            total += (size_t)nq * lnsib[li] * 4;
        }
    }
    const size_t exchanged_words = total;
    DevMem d_out(total);
    if (ctx->world > 1) OLA_CUDA(cudaMemsetAsync(d_out.p, 0, total * 8, st));
    bool sharded = false;
    for (size_t o = 0; o < oracles.size(); ++o) {
        size_t cnt = (size_t)nq * oracles[o]->ncols;
        // a coset shard answers the queries that fall into its leaf range; the others come from their owners below
        const size_t Lo = (size_t)1 << oracles[o]->leaf_bits(), leaf_lo = (size_t)oracles[o]->coset_first * n;
        sharded = sharded || Lo != L;
        {
            Launch lz(ctx, "fri_query_rows");
            gather_query_rows_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(oracles[o]->d_lde, Lo, oracles[o]->ncols, (const uint32_t*)d_idx.p, 0, nq,
                                                                                 leaf_lo, d_out.p + off_rows[o]);
        }
        if (nsib0 > 0) {
            Launch lz(ctx, "fri_query_paths");
            size_t pc = (size_t)nq * nsib0 * 4;
            gather_query_paths_kernel<<<(unsigned)((pc + 127) / 128), 128, 0, st>>>(oracles[o]->d_nodes, Lo, (const uint32_t*)d_idx.p, 0, nq, nsib0,
                                                                                  leaf_lo, d_out.p + off_paths[o]);
        }
    }
    for (size_t li = 0; li < layers.size(); ++li) {
        const int arity = 1 << arities[li];
        size_t cnt = (size_t)nq * 2 * arity;
        // the owner of the leaf answers; a replicated layer is answered by rank 0 alone when the buffer is exchanged
        const bool answer = layers[li]->sharded || !sharded || ctx->rank == 0;
        if (!answer) continue;
        {
            Launch lz(ctx, "fri_query_rows");
            gather_query_ext_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(layers[li]->vals.p, layers[li]->len_loc, arity, (const uint32_t*)d_idx.p, lshift[li],
                                                                                nq, layers[li]->leaf_lo, layers[li]->nloc, d_out.p + off_lrows[li]);
        }
        if (lnsib[li] > 0) {
            Launch lz(ctx, "fri_query_paths");
            size_t pc = (size_t)nq * lnsib[li] * 4;
            gather_query_paths_kernel<<<(unsigned)((pc + 127) / 128), 128, 0, st>>>(layers[li]->nodes.p, layers[li]->nloc, (const uint32_t*)d_idx.p, lshift[li], nq,
                                                                                  lnsib[li], layers[li]->leaf_lo, d_out.p + off_lpaths[li]);
        }
    }
    if (sharded) {
        // owner-answered openings: every rank holds zeros for the leaves it does not own -> wrapping sum = the answer
        comm_allreduce(ctx, d_out.p, exchanged_words);
    }
    check_launch("fri query gathers");
    std::vector<uint64_t> hout(total);
    OLA_CUDA(cudaMemcpyAsync(hout.data(), d_out.p, total * 8, cudaMemcpyDeviceToHost, st));
    OLA_CUDA(cudaStreamSynchronize(st));
    auto hashes = [&](size_t off, int count) {
        std::vector<stark::Hash> v(count);
        memcpy(v.data(), &hout[off], (size_t)count * 32);
        return v;
    };
    for (int q = 0; q < nq; ++q) {
        stark::FriQueryRound qr;
        for (size_t o = 0; o < oracles.size(); ++o) {
            size_t nc = oracles[o]->ncols;
            std::vector<F> row(hout.begin() + off_rows[o] + (size_t)q * nc, hout.begin() + off_rows[o] + (size_t)(q + 1) * nc);
            qr.initial.push_back({row, hashes(off_paths[o] + (size_t)q * nsib0 * 4, nsib0)});
        }
        for (size_t li = 0; li < layers.size(); ++li) {
            const int arity = 1 << arities[li];
            stark::FriQueryStep stp;
            const uint64_t* r = &hout[off_lrows[li] + (size_t)q * 2 * arity];
            for (int e = 0; e < arity; ++e) stp.evals.push_back(gl::make2(r[2 * e], r[2 * e + 1]));
            stp.siblings = hashes(off_lpaths[li] + (size_t)q * lnsib[li] * 4, lnsib[li]);
            qr.steps.push_back(std::move(stp));
        }
        proof.rounds.push_back(std::move(qr));
    }
    return proof;
}

}  // namespace fri
}  // namespace ola
