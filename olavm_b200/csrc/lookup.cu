// Lookup-argument columns and the RangeCheck table on the GPU (trace generation next to the prover).
//
// Replaces
//   circuits/src/stark/lookup.rs:68-131          permuted_cols: sort the input and the table column, then ONE sequential merge
//                                                walk with a LIFO list of skipped table values and a FIFO list of unfilled rows
//   circuits/src/generation/builtin.rs:249-316   generate_rc_trace (two permuted_cols calls over 2^16 .. 2^22+ rows)
//
// The walk is serial in the reference; here it is re-derived as data-parallel steps that produce the SAME columns:
//   1. S = sort(canonical inputs), T = sort(canonical table)                   (bitonic network, shared-memory inner stages)
//   2. input i (value v, rank r among the equal inputs) is paired with a table copy of v iff r < #{T == v}: permuted_table[i] = v.
//      Table entry j (value w, rank r') is paired iff r' < #{S == w}.                                   (binary searches)
//   3. The unpaired entries are the walk's events in increasing value: an unpaired table entry is PUSHED on the list of unused
//      values, an unpaired input POPS the most recent one (pushes and pops never share a value: one side has the surplus).  The
//      walk stops when the table is exhausted, so an unpaired input whose value is >= max(T) pops nothing.  The position of an
//      event in that order needs no sort: (#pushes before it in T) + (#pops before it in S), two prefix sums.
//   4. "pop takes the most recent unused push, a pop on an empty list waits" is bracket matching: with depth = running sum of
//      +1 / -1, a push at depth d (after) and a pop at depth d (before) alternate inside every level, so after sorting the events
//      by (level, position) each pop is matched by its predecessor iff that is a push of the same level.
//   5. The unmatched inputs (in row order) receive the never-popped table values (in table order): two compactions.
// Every step is exact integer work; tests compare with the oracle's statement-by-statement walk on valid lookups, missing
// values, surplus on either side, duplicates in the table and random 64-bit columns.
#include "batch.h"
#include "common.h"
#include "gl.cuh"
#include "stark_types.h"

#include <vector>

namespace ola {
namespace lookup {

static constexpr int SORT_TILE = 2048;  // elements sorted per CTA in shared memory (1024 threads, two each)

// ---- bitonic sort of 2^k u64 keys, ascending ----------------------------------------------------------------------------------
__device__ __forceinline__ void cmp_swap(uint64_t& a, uint64_t& b, bool asc) {
    if ((a > b) == asc) {
        const uint64_t t = a;
        a = b;
        b = t;
    }
}
// all stages k = 2 .. min(n, SORT_TILE) inside one tile
__global__ void __launch_bounds__(SORT_TILE / 2) sort_tile_kernel(uint64_t* __restrict__ d, size_t n) {
    __shared__ uint64_t s[SORT_TILE];
    const size_t base = (size_t)blockIdx.x * SORT_TILE;
    const int t = threadIdx.x;
    const int m = (int)(n < (size_t)SORT_TILE ? n : (size_t)SORT_TILE);
    for (int i = t; i < m; i += blockDim.x) s[i] = d[base + i];
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = t; p < m / 2; p += blockDim.x) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));  // the lower index of the pair
                const bool asc = (((base + i) & (size_t)k) == 0);
                cmp_swap(s[i], s[i + j], asc);
            }
            __syncthreads();
        }
    }
    for (int i = t; i < m; i += blockDim.x) d[base + i] = s[i];
}
// one global step (distance j >= SORT_TILE) of stage k
__global__ void sort_global_step_kernel(uint64_t* __restrict__ d, size_t n, size_t k, size_t j) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n / 2) return;
    const size_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
    uint64_t a = d[i], b = d[i + j];
    const bool asc = ((i & k) == 0);
    if ((a > b) == asc) {
        d[i] = b;
        d[i + j] = a;
    }
}
// the steps j = SORT_TILE / 2 .. 1 of stage k (k > SORT_TILE) inside one tile
__global__ void __launch_bounds__(SORT_TILE / 2) sort_tile_tail_kernel(uint64_t* __restrict__ d, size_t k) {
    __shared__ uint64_t s[SORT_TILE];
    const size_t base = (size_t)blockIdx.x * SORT_TILE;
    const int t = threadIdx.x;
    for (int i = t; i < SORT_TILE; i += blockDim.x) s[i] = d[base + i];
    __syncthreads();
    const bool asc = ((base & k) == 0);  // k > SORT_TILE: one direction for the whole tile
    for (int j = SORT_TILE >> 1; j > 0; j >>= 1) {
        for (int p = t; p < SORT_TILE / 2; p += blockDim.x) {
            const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
            cmp_swap(s[i], s[i + j], asc);
        }
        __syncthreads();
    }
    for (int i = t; i < SORT_TILE; i += blockDim.x) d[base + i] = s[i];
}
static void sort_u64(ola_ctx* ctx, uint64_t* d, size_t n) {  // n a power of two
    Launch lz(ctx, "lookup_sort");
    const size_t tiles = n <= (size_t)SORT_TILE ? 1 : n / SORT_TILE;
    sort_tile_kernel<<<(unsigned)tiles, SORT_TILE / 2, 0, ctx->stream>>>(d, n);
    for (size_t k = (size_t)SORT_TILE * 2; k <= n; k <<= 1) {
        for (size_t j = k >> 1; j >= (size_t)SORT_TILE; j >>= 1)
            sort_global_step_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, ctx->stream>>>(d, n, k, j);
        sort_tile_tail_kernel<<<(unsigned)tiles, SORT_TILE / 2, 0, ctx->stream>>>(d, k);
    }
    check_launch("lookup sort");
}

// ---- exclusive prefix sums of 32-bit counts (signed values wrap correctly) ---------------------------------------------------------
static constexpr int SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(SCAN_BLOCK) scan_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ sums, size_t n) {
    __shared__ uint32_t s[SCAN_BLOCK];
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
        const uint32_t a = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += a;
        __syncthreads();
    }
    if (i < n) out[i] = s[threadIdx.x] - v;  // exclusive
    if (threadIdx.x == SCAN_BLOCK - 1) sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(uint32_t* __restrict__ sums, size_t nblocks, uint32_t* __restrict__ total) {
    // one CTA walks the block sums in chunks of SCAN_BLOCK (at most 2^24 / 2^10 = 16384 of them)
    __shared__ uint32_t s[SCAN_BLOCK];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nblocks; base += SCAN_BLOCK) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < SCAN_BLOCK; off <<= 1) {
            const uint32_t a = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
            __syncthreads();
            s[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < nblocks) sums[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ sums, size_t n) {
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += sums[blockIdx.x];
}
// out[i] = sum_{k < i} in[k];  *total (device) = the sum of all
static void exclusive_scan(ola_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t* sums, uint32_t* total, size_t n) {
    const size_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scan_block_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(in, out, sums, n);
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, ctx->stream>>>(sums, nb, total);
    scan_add_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(out, sums, n);
    check_launch("lookup scan");
}

// ---- steps 2-5 ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t lower_bound(const uint64_t* __restrict__ a, size_t n, uint64_t v) {  // first index with a[i] >= v
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (a[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
__device__ __forceinline__ size_t upper_bound(const uint64_t* __restrict__ a, size_t n, uint64_t v) {  // first index with a[i] > v
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (a[mid] <= v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void canon_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gl::canon(in[i]);
}

// step 2: pairing.  pop[i] = 1: input i is unpaired and its value is below max(T) (it pops);  hole[i] = 1: unpaired at or above
// max(T) (nothing to pop: the walk has ended).  push[j] = 1: table entry j is unpaired.  Paired inputs get their table value.
__global__ void classify_kernel(const uint64_t* __restrict__ S, const uint64_t* __restrict__ T, size_t n, uint64_t* __restrict__ PT,
                                uint32_t* __restrict__ pop, uint32_t* __restrict__ hole, uint32_t* __restrict__ push) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t tmax = T[n - 1];
    {
        const uint64_t v = S[i];
        const size_t r = i - lower_bound(S, n, v);
        const size_t cnt_t = upper_bound(T, n, v) - lower_bound(T, n, v);
        const bool paired = r < cnt_t;
        PT[i] = paired ? v : 0;
        pop[i] = (!paired && v < tmax) ? 1u : 0u;
        hole[i] = (!paired && v >= tmax) ? 1u : 0u;
    }
    {
        const uint64_t w = T[i];
        const size_t r = i - lower_bound(T, n, w);
        const size_t cnt_s = upper_bound(S, n, w) - lower_bound(S, n, w);
        push[i] = (r < cnt_s) ? 0u : 1u;
    }
}
// step 3: the events in walk order.  ev_delta[e] = +1 (push) / -1 (pop), ev_ref[e] = j (push) or n + i (pop)
__global__ void events_kernel(const uint64_t* __restrict__ S, const uint64_t* __restrict__ T, size_t n, const uint32_t* __restrict__ pop,
                              const uint32_t* __restrict__ push, const uint32_t* __restrict__ cum_pop, const uint32_t* __restrict__ cum_push,
                              uint32_t total_pop, uint32_t total_push, uint32_t* __restrict__ ev_delta, uint32_t* __restrict__ ev_ref) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pop[i]) {  // pushes of smaller values come first (none has the same value)
        const size_t lb = lower_bound(T, n, S[i]);
        const uint32_t e = cum_pop[i] + (lb < n ? cum_push[lb] : total_push);
        ev_delta[e] = 0xFFFFFFFFu;
        ev_ref[e] = (uint32_t)(n + i);
    }
    if (push[i]) {
        const size_t lb = lower_bound(S, n, T[i]);
        const uint32_t e = cum_push[i] + (lb < n ? cum_pop[lb] : total_pop);
        ev_delta[e] = 1u;
        ev_ref[e] = (uint32_t)i;
    }
}
// step 4a: sort keys (level, position).  depth_before[e] = exclusive sum of the deltas; a push sits at level depth_before + 1,
// a pop at level depth_before; levels are offset by n_events so that they stay non-negative
__global__ void level_keys_kernel(const uint32_t* __restrict__ ev_delta, const uint32_t* __restrict__ depth_before, uint32_t n_events, size_t cap,
                                  uint64_t* __restrict__ keys) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cap) return;
    if (e >= n_events) {
        keys[e] = ~0ULL;  // padding sorts last
        return;
    }
    const int32_t level = (int32_t)depth_before[e] + (ev_delta[e] == 1u ? 1 : 0);
    keys[e] = ((uint64_t)(uint32_t)(level + (int32_t)n_events) << 32) | (uint64_t)e;
}
// step 4b: a pop whose predecessor in (level, position) order is a push of the same level takes that push's table value
__global__ void match_kernel(const uint64_t* __restrict__ keys, uint32_t n_events, const uint32_t* __restrict__ ev_delta, const uint32_t* __restrict__ ev_ref,
                             const uint64_t* __restrict__ T, size_t n, uint64_t* __restrict__ PT, uint32_t* __restrict__ hole, uint32_t* __restrict__ left) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_events) return;
    const uint64_t key = keys[s];
    const uint32_t e = (uint32_t)key;
    if (ev_delta[e] == 1u) {  // a push: it stays on the list unless the next event of its level pops it
        bool popped = false;
        if (s + 1 < n_events) {
            const uint64_t nk = keys[s + 1];
            popped = (nk >> 32) == (key >> 32) && ev_delta[(uint32_t)nk] != 1u;
        }
        left[ev_ref[e]] = popped ? 0u : 1u;
        return;
    }
    const uint32_t i = ev_ref[e] - (uint32_t)n;
    bool matched = false;
    if (s > 0) {
        const uint64_t pk = keys[s - 1];
        if ((pk >> 32) == (key >> 32) && ev_delta[(uint32_t)pk] == 1u) {
            PT[i] = T[ev_ref[(uint32_t)pk]];
            matched = true;
        }
    }
    if (!matched) hole[i] = 1u;  // popped an empty list: filled at the end
}
// step 5: the k-th unfilled row takes the k-th value left on the list
__global__ void compact_left_kernel(const uint64_t* __restrict__ T, size_t n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ cum_left,
                                    uint64_t* __restrict__ left_vals) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && left[j]) left_vals[cum_left[j]] = T[j];
}
__global__ void fill_holes_kernel(size_t n, const uint32_t* __restrict__ hole, const uint32_t* __restrict__ cum_hole, const uint64_t* __restrict__ left_vals,
                                  uint64_t* __restrict__ PT) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && hole[i]) PT[i] = left_vals[cum_hole[i]];
}
__global__ void zero_u32_kernel(uint32_t* p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

struct Buf {
    uint64_t* p = nullptr;
    explicit Buf(size_t n_u64) { dev_alloc(&p, n_u64 ? n_u64 : 1); }
    ~Buf() {
        if (p) dev_free(p);
    }
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    uint32_t* u32() { return reinterpret_cast<uint32_t*>(p); }
};

// d_inputs, d_table: n elements each (any representatives); d_perm_inputs, d_perm_table: n each.  n a power of two >= 2.
void permuted_cols(ola_ctx* ctx, const uint64_t* d_inputs, const uint64_t* d_table, size_t n, uint64_t* d_perm_inputs, uint64_t* d_perm_table) {
    OLA_CHECK(n >= 2 && (n & (n - 1)) == 0 && n <= ((size_t)1 << 24), OLA_ERR_INVALID_ARG, "permuted_cols: the column length must be a power of two in [2, 2^24]");
    const unsigned gb = (unsigned)((n + 255) / 256);
    uint64_t* S = d_perm_inputs;  // the sorted inputs ARE the permuted inputs (lookup.rs:130)
    Buf T(n);
    canon_kernel<<<gb, 256, 0, ctx->stream>>>(d_inputs, S, n);
    canon_kernel<<<gb, 256, 0, ctx->stream>>>(d_table, T.p, n);
    sort_u64(ctx, S, n);
    sort_u64(ctx, T.p, n);

    Launch lz(ctx, "lookup_walk");
    // 32-bit work arrays: pop, hole, push, left | cum_pop, cum_push, cum_x | block sums | totals
    const size_t nb = (2 * n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    Buf w32((7 * n + nb + 16) / 2 + 8);
    uint32_t* pop = w32.u32();
    uint32_t* hole = pop + n;
    uint32_t* push = hole + n;
    uint32_t* left = push + n;
    uint32_t* cum_a = left + n;
    uint32_t* cum_b = cum_a + n;
    uint32_t* cum_c = cum_b + n;
    uint32_t* sums = cum_c + n;
    uint32_t* totals = sums + nb;  // [0] pops, [1] pushes, [2] scratch
    classify_kernel<<<gb, 256, 0, ctx->stream>>>(S, T.p, n, d_perm_table, pop, hole, push);
    exclusive_scan(ctx, pop, cum_a, sums, totals + 0, n);
    exclusive_scan(ctx, push, cum_b, sums, totals + 1, n);
    uint32_t h_tot[2];
    OLA_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(h_tot), cudaMemcpyDeviceToHost, ctx->stream));
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t n_pop = h_tot[0], n_push = h_tot[1], n_ev = n_pop + n_push;
    zero_u32_kernel<<<gb, 256, 0, ctx->stream>>>(left, n);
    if (n_ev) {
        size_t cap = 2;
        while (cap < n_ev) cap <<= 1;
        Buf ev(cap + cap + cap);  // ev_delta | ev_ref | depth (u32 each, cap entries) + keys (u64, cap entries)
        uint32_t* ev_delta = ev.u32();
        uint32_t* ev_ref = ev_delta + cap;
        uint32_t* depth = ev_ref + cap;
        uint64_t* keys = ev.p + (3 * cap + 1) / 2 + 1;
        Buf ev_sums((cap / SCAN_BLOCK + 2) / 2 + 2);
        events_kernel<<<gb, 256, 0, ctx->stream>>>(S, T.p, n, pop, push, cum_a, cum_b, n_pop, n_push, ev_delta, ev_ref);
        exclusive_scan(ctx, ev_delta, depth, ev_sums.u32(), totals + 2, n_ev);
        level_keys_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(ev_delta, depth, n_ev, cap, keys);
        sort_u64(ctx, keys, cap);
        match_kernel<<<(unsigned)((n_ev + 255) / 256), 256, 0, ctx->stream>>>(keys, n_ev, ev_delta, ev_ref, T.p, n, d_perm_table, hole, left);
        check_launch("lookup match");
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));  // ev / ev_sums are released below: the stream-ordered pool makes that safe, the sync keeps errors local
    }
    // a table entry that was never an event is paired; one that was pushed and never popped is left over
    exclusive_scan(ctx, left, cum_c, sums, totals + 2, n);
    Buf left_vals(n);
    compact_left_kernel<<<gb, 256, 0, ctx->stream>>>(T.p, n, left, cum_c, left_vals.p);
    exclusive_scan(ctx, hole, cum_a, sums, totals + 2, n);
    fill_holes_kernel<<<gb, 256, 0, ctx->stream>>>(n, hole, cum_a, left_vals.p, d_perm_table);
    check_launch("lookup fill");
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- generate_rc_trace (builtin.rs:249-316) ---------------------------------------------------------------------------------------------
// vals[nrows], kinds[nrows] in {0 cpu, 1 memory sort, 2 memory region, 3 comparison} (one u64 each); out = [12][n] column-major
__global__ void rc_fill_kernel(const uint64_t* __restrict__ vals, const uint64_t* __restrict__ kinds, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool live = i < nrows;
    const uint64_t v = live ? gl::canon(vals[i]) : 0;
    const uint64_t k = live ? kinds[i] : 4;
    out[0 * n + i] = k == 0;          // CPU_FILTER
    out[1 * n + i] = k == 1;          // MEMORY_SORT_FILTER
    out[2 * n + i] = k == 2;          // MEMORY_REGION_FILTER
    out[3 * n + i] = k == 3;          // CMP_FILTER
    out[4 * n + i] = v;               // VAL
    out[5 * n + i] = v & 0xFFFF;      // LIMB_LO  (split_u16_limbs_from_field, trace.rs:414-418)
    out[6 * n + i] = v >> 16;         // LIMB_HI
    out[9 * n + i] = i < 65536 ? i : 65535;  // FIX_RANGE_CHECK_U16, padded with its last value (builtin.rs:278-293)
}
void rangecheck_trace(ola_ctx* ctx, const uint64_t* d_vals, const uint64_t* d_kinds, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    OLA_CHECK(log_n >= 16 && nrows <= n, OLA_ERR_INVALID_ARG, "RangeCheck table: at least 2^16 rows (RANGE_CHECK_U16_SIZE) and room for every value");
    {
        Launch lz(ctx, "gen_rc_fill");
        rc_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_vals, d_kinds, nrows, n, d_out);
        check_launch("rc_fill_kernel");
    }
    permuted_cols(ctx, d_out + 5 * n, d_out + 9 * n, n, d_out + 7 * n, d_out + 10 * n);  // LIMB_LO_PERMUTED, FIX_..._PERMUTED_LO
    permuted_cols(ctx, d_out + 6 * n, d_out + 9 * n, n, d_out + 8 * n, d_out + 11 * n);  // LIMB_HI_PERMUTED, FIX_..._PERMUTED_HI
}

// ---- generate_bitwise_trace (builtin.rs:35-206) --------------------------------------------------------------------------------------
// Column order builtins/bitwise/columns.rs:23-62 (59 columns).  The reference writes the FOURTH limb of op0 / op1 / res to
// trace[OP0_LIMBS.end] / [OP1_LIMBS.end] / [RES_LIMBS.end] (builtin.rs:66, :71, :76) -- the first column of the NEXT range,
// overwritten right after -- so columns 8, 12 and 16 stay zero in the table it produces; a drop-in generator reproduces that
// table (operands below 2^24 satisfy the AIR, as in the reference).  Opcode::AND = 18, OR = 17, XOR = 16 (instruction.rs:44-46).
__global__ void bitwise_fill_kernel(const uint64_t* __restrict__ tags, const uint64_t* __restrict__ op0, const uint64_t* __restrict__ op1,
                                    const uint64_t* __restrict__ res, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto put = [&](int c, uint64_t v) { out[(size_t)c * n + i] = v; };
    const bool live = i < nrows;
    const uint64_t a = live ? gl::canon(op0[i]) : 0, b = live ? gl::canon(op1[i]) : 0, r = live ? gl::canon(res[i]) : 0;
    put(0, live ? 1 : 0);
    put(1, live ? tags[i] : 0);
    put(2, a);
    put(3, b);
    put(4, r);
    put(5, a & 255), put(6, (a >> 8) & 255), put(7, (a >> 16) & 255), put(8, 0);
    put(9, b & 255), put(10, (b >> 8) & 255), put(11, (b >> 16) & 255), put(12, 0);
    put(13, r & 255), put(14, (r >> 8) & 255), put(15, (r >> 16) & 255), put(16, 0);
    put(37, i < 256 ? i : 0);  // FIX_RANGE_CHECK_U8: 0 .. 255, then zeros (builtin.rs:85; not padded with its last value)
    const size_t blk = i >> 16, idx = i & 65535, x = idx >> 8, y = idx & 255;
    const bool fixed = blk < 3;
    put(50, fixed ? (1ull << (18 - blk)) : 0);  // FIX_TAG: AND, OR, XOR blocks of 2^16 rows
    put(51, fixed ? x : 0);
    put(52, fixed ? y : 0);
    put(53, fixed ? (blk == 0 ? (x & y) : blk == 1 ? (x | y) : (x ^ y)) : 0);
}
// COMPRESS_LIMBS[k] = tag + op0_k beta + op1_k beta^2 + res_k beta^3;  FIX_COMPRESS likewise over the fixed columns (builtin.rs:133-159)
__global__ void bitwise_compress_kernel(size_t n, uint64_t beta, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t b2 = gl::mul(beta, beta), b3 = gl::mul(b2, beta);
    auto at = [&](int c) { return out[(size_t)c * n + i]; };
    const uint64_t tag = gl::canon(at(1));
#pragma unroll
    for (int k = 0; k < 4; ++k)
        out[(size_t)(29 + k) * n + i] = gl::add(gl::add(gl::add(tag, gl::mul(at(5 + k), beta)), gl::mul(at(9 + k), b2)), gl::mul(at(13 + k), b3));
    out[(size_t)54 * n + i] = gl::add(gl::add(gl::add(at(50), gl::mul(at(51), beta)), gl::mul(at(52), b2)), gl::mul(at(53), b3));
}
// The generator in three steps, so that a caller can run the middle one -- host-only, sequential, 3 * n / 2 permutations -- on
// another thread while the GPU does other work (ola_prove_trace commits the other eleven tables meanwhile):
//   bitwise_trace_begin   row fill and the twelve lookups that do not involve beta; returns the twelve limb columns (host)
//   bitwise_beta          the compress challenge: a fresh Challenger observes the limb columns (builtin.rs:118-131); no CUDA
//   bitwise_trace_finish  the beta-compressed columns and their four lookups
// d_out = [59][n]
std::vector<uint64_t> bitwise_trace_begin(ola_ctx* ctx, const uint64_t* d_tags, const uint64_t* d_op0, const uint64_t* d_op1, const uint64_t* d_res,
                                          size_t nrows, uint32_t log_n, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    OLA_CHECK(log_n >= 18 && log_n <= 24 && nrows <= n, OLA_ERR_INVALID_ARG, "Bitwise table: at least 2^18 rows (3 * 2^16 fixed rows) and room for every operation");
    const unsigned gb = (unsigned)((n + 255) / 256);
    {
        Launch lz(ctx, "gen_bitwise_fill");
        bitwise_fill_kernel<<<gb, 256, 0, ctx->stream>>>(d_tags, d_op0, d_op1, d_res, nrows, n, d_out);
        check_launch("bitwise_fill_kernel");
    }
    std::vector<uint64_t> limbs(12 * n);
    OLA_CUDA(cudaMemcpyAsync(limbs.data(), d_out + 5 * n, 12 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    auto col = [&](int c) { return d_out + (size_t)c * n; };
    for (int k = 0; k < 4; ++k) {  // builtin.rs:162-195, the byte range checks
        permuted_cols(ctx, col(5 + k), col(37), n, col(17 + k), col(38 + k));
        permuted_cols(ctx, col(9 + k), col(37), n, col(21 + k), col(38 + 4 + k));
        permuted_cols(ctx, col(13 + k), col(37), n, col(25 + k), col(38 + 8 + k));
    }
    OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    return limbs;
}
uint64_t bitwise_beta(const std::vector<uint64_t>& limbs) {
    stark::Challenger ch(OLA_HASH_POSEIDON);
    for (size_t k = 0; k < limbs.size(); ++k) ch.observe(limbs[k]);
    return ch.get_challenge();
}
void bitwise_trace_finish(ola_ctx* ctx, uint32_t log_n, uint64_t beta, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    {
        Launch lz(ctx, "gen_bitwise_compress");
        bitwise_compress_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, beta, d_out);
        check_launch("bitwise_compress_kernel");
    }
    auto col = [&](int c) { return d_out + (size_t)c * n; };
    for (int k = 0; k < 4; ++k) permuted_cols(ctx, col(29 + k), col(54), n, col(33 + k), col(55 + k));  // the compressed triples against FIX_COMPRESS
}
// returns beta
uint64_t bitwise_trace(ola_ctx* ctx, const uint64_t* d_tags, const uint64_t* d_op0, const uint64_t* d_op1, const uint64_t* d_res, size_t nrows, uint32_t log_n,
                       uint64_t* d_out) {
    const uint64_t beta = bitwise_beta(bitwise_trace_begin(ctx, d_tags, d_op0, d_op1, d_res, nrows, log_n, d_out));
    bitwise_trace_finish(ctx, log_n, beta, d_out);
    return beta;
}

// ---- generate_cmp_trace (builtin.rs:208-247) ------------------------------------------------------------------------------------------
// cells [nrows][6] = (op0, op1, gte, abs_diff, abs_diff_inv, filter_looking_rc) as the executor recorded them; padding rows
// (1, 0, 1, 1, 1, 0).  d_out = [6][n]
__global__ void cmp_fill_kernel(const uint64_t* __restrict__ cells, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i < nrows) {
#pragma unroll
        for (int c = 0; c < 6; ++c) out[(size_t)c * n + i] = gl::canon(cells[i * 6 + c]);
    } else {
        out[0 * n + i] = 1, out[1 * n + i] = 0, out[2 * n + i] = 1, out[3 * n + i] = 1, out[4 * n + i] = 1, out[5 * n + i] = 0;
    }
}
void cmp_trace(ola_ctx* ctx, const uint64_t* d_cells, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    OLA_CHECK(log_n >= 1 && nrows <= n, OLA_ERR_INVALID_ARG, "Cmp table: room for every comparison");
    Launch lz(ctx, "gen_cmp_fill");
    cmp_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_cells, nrows, n, d_out);
    check_launch("cmp_fill_kernel");
}

// ---- generate_cpu_trace (circuits/src/generation/cpu.rs:11-218) -------------------------------------------------------------------------
// One thread per table row: the executor's Step record (66 u64, layout in include/ola_gpu.h) is copied into the 94 columns
// (cpu/columns.rs), the opcode picks its selector column (opcode_to_selector, cpu.rs:19-59), the derived flags follow
// cpu.rs:111-177 on the raw u64 values, rows past nrows are padding rows (cpu.rs:180-208: END opcode, the last row's
// instruction and storage index carried on).
__device__ __forceinline__ int cpu_selector_of(uint64_t opcode) {
    // binary_bit_mask = 1 << shift (core/src/vm/opcodes.rs): ADD 31, MUL 30, EQ 29, ASSERT 28, MOV 27, JMP 26, CJMP 25, CALL 24, RET 23,
    // MLOAD 22, MSTORE 21, END 20, RC 19, AND 18, OR 17, XOR 16, NOT 15, NEQ 14, GTE 13, POSEIDON 12, SLOAD 11, SSTORE 10, TLOAD 9,
    // TSTORE 8, SCCALL 7
    if (opcode == 0 || (opcode & (opcode - 1)) != 0) return -1;
    const int sh = 63 - __clzll((long long)opcode);
    switch (sh) {
        case 31: case 30: case 29: case 28: case 14: return 66;  // COL_S_SIMPLE_ARITHMATIC_OP
        case 27: return 67;
        case 26: return 68;
        case 25: return 69;
        case 24: return 70;
        case 23: return 71;
        case 22: return 72;
        case 21: return 73;
        case 20: return 74;
        case 19: return 75;
        case 18: case 17: case 16: return 76;  // COL_S_BITWISE
        case 15: return 77;
        case 13: return 78;
        case 12: return 79;
        case 11: return 80;
        case 10: return 81;
        case 9: return 82;
        case 8: return 83;
        case 7: return 84;
        default: return -1;
    }
}
__global__ void cpu_fill_kernel(const uint64_t* __restrict__ steps /* [nrows][66] */, size_t nrows, size_t n, uint64_t* __restrict__ out /* [94][n] */) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto put = [&](int c, uint64_t v) { out[(size_t)c * n + i] = v; };
    constexpr uint64_t END = 1ull << 20, SLOAD = 1ull << 11, SSTORE = 1ull << 10, TLOAD = 1ull << 9, TSTORE = 1ull << 8, SCCALL = 1ull << 7,
                       MLOAD = 1ull << 22, MSTORE = 1ull << 21;
    if (i >= nrows) {  // padding
        const uint64_t* last = nrows ? steps + (nrows - 1) * 66 : nullptr;
        for (int c = 0; c < 94; ++c) put(c, 0);
        put(26, last ? gl::canon(last[25]) : 1048576);  // COL_INST of the last executed row
        put(28, END);
        put(35, last ? gl::canon(last[34]) : 0);        // COL_IDX_STORAGE carried on
        put(74, 1), put(85, 1), put(86, 1), put(93, 1);
        return;
    }
    const uint64_t* s = steps + i * 66;
    put(0, 0);
    put(1, gl::canon(s[0])), put(2, gl::canon(s[1]));
    for (int j = 0; j < 4; ++j) put(3 + j, gl::canon(s[2 + j])), put(7 + j, gl::canon(s[6 + j]));
    put(11, gl::canon(s[10]));
    put(12, (uint32_t)s[11]);
    put(13, gl::canon(s[12])), put(14, gl::canon(s[13])), put(15, gl::canon(s[14]));
    for (int j = 0; j < 10; ++j) put(16 + j, gl::canon(s[15 + j]));
    for (int j = 0; j < 10; ++j) put(26 + j, gl::canon(s[25 + j]));  // inst, op1_imm, opcode, imm, op0, op1, dst, aux0, aux1, idx_storage
    for (int j = 0; j < 30; ++j) put(36 + j, gl::canon(s[35 + j]));  // the three register-selector groups
    const uint64_t opcode = s[27], op0 = s[29], op1 = s[30], ext_cnt = s[14], is_ext = s[13];
    const bool env_zero = gl::canon(s[0]) == 0;
    const int sel = cpu_selector_of(opcode);
    for (int c = 66; c < 85; ++c) put(c, c == sel ? 1 : 0);
    put(85, env_zero ? 1 : 0);
    uint64_t ext_length = 0;  // cpu.rs:118-132: plain u64 arithmetic on the inner values
    if (opcode == SLOAD || opcode == SSTORE || opcode == SCCALL || (opcode == END && !env_zero))
        ext_length = 1;
    else if (opcode == TLOAD)
        ext_length = op0 * op1 + (1 - op0);
    else if (opcode == TSTORE)
        ext_length = op1;
    put(86, ext_length == ext_cnt ? 1 : 0);
    put(87, (env_zero && opcode == END) ? 0 : 1);
    put(88, gl::canon(s[65]));
    put(89, (opcode == SCCALL && ext_cnt == 1) ? 1 : 0);
    put(90, ((opcode == SLOAD || opcode == SSTORE) && is_ext == 1) ? 1 : 0);
    put(91, (opcode == END && is_ext == 1) ? 1 : 0);
    put(92, is_ext == 1 ? 0 : ((opcode == MLOAD || opcode == MSTORE) ? 1 : (s[26] == 1 ? 1 : 0)));
    put(93, 0);
}
void cpu_trace(ola_ctx* ctx, const uint64_t* d_steps, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    OLA_CHECK(nrows <= n, OLA_ERR_INVALID_ARG, "CPU table: room for every executed row");
    Launch lz(ctx, "gen_cpu_fill");
    cpu_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_steps, nrows, n, d_out);
    check_launch("cpu_fill_kernel");
}

// ---- generate_memory_trace (circuits/src/generation/memory.rs:8-155) -------------------------------------------------------------------
// One thread per table row.  Filled rows copy the executor's MemoryTraceCell (15 u64, layout in include/ola_gpu.h; already
// sorted and differenced by gen_memory_table), pick the selector of the accessing opcode and derive the two range-check
// filters (the first needs the previous cell's region); padding rows continue the write-once region one address per row
// (memory.rs:113-146), the first of them carrying the inverse of its distance to the last filled row.
__device__ __forceinline__ int mem_selector_of(uint64_t op) {
    if (op == 0) return 16;  // COL_MEM_S_PROPHET
    if (op & (op - 1)) return -1;
    switch (63 - __clzll((long long)op)) {
        case 22: return 6;   // MLOAD
        case 21: return 7;   // MSTORE
        case 24: return 8;   // CALL
        case 23: return 9;   // RET
        case 9: return 10;   // TLOAD
        case 8: return 11;   // TSTORE
        case 7: return 12;   // SCCALL
        case 12: return 13;  // POSEIDON
        case 10: return 14;  // SSTORE
        case 11: return 15;  // SLOAD
        default: return -1;
    }
}
__global__ void memory_fill_kernel(const uint64_t* __restrict__ cells /* [ncells][15] */, size_t ncells, size_t n, uint64_t* __restrict__ out /* [29][n] */) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto put = [&](int c, uint64_t v) { out[(size_t)c * n + i] = v; };
    constexpr uint64_t SPAN = 0xFFFFFFFFull;  // 2^32 - 1
    if (i < ncells) {
        const uint64_t* c = cells + i * 15;
        put(0, 0);
        for (int j = 0; j < 5; ++j) put(1 + j, gl::canon(c[j]));  // env_idx, is_rw, addr, clk, op
        const int sel = mem_selector_of(c[4]);
        for (int k = 6; k <= 16; ++k) put(k, k == sel ? 1 : 0);
        for (int j = 0; j < 10; ++j) put(17 + j, gl::canon(c[5 + j]));  // is_write, value, diff_addr, its inverse, diff_clk, cond, unchanged, prophet, heap, rc_value
        const bool prophet = gl::canon(c[12]) == 1, heap = gl::canon(c[13]) == 1;
        const bool last_is_not_heap = i > 0 && gl::canon(cells[(i - 1) * 15 + 13]) == 0;
        put(27, (i == 0 || prophet || (heap && last_is_not_heap)) ? 0 : 1);
        put(28, (heap || prophet) ? 1 : 0);
        return;
    }
    // no cell at all: row 0 is the first address of the write-once region (memory.rs:101-112) and counts as filled
    const size_t filled = ncells ? ncells : 1;
    const uint64_t last_is_rw = ncells ? gl::canon(cells[(ncells - 1) * 15 + 1]) : 0;
    const uint64_t last_addr = ncells ? gl::canon(cells[(ncells - 1) * 15 + 2]) : gl::sub(0, SPAN);
    const uint64_t last_env = ncells ? gl::canon(cells[(ncells - 1) * 15 + 0]) : 0;
    for (int k = 0; k < 29; ++k) put(k, 0);
    if (i == 0) {  // only when ncells == 0
        put(3, last_addr), put(17, 1), put(22, gl::sub(0, last_addr)), put(24, 1), put(26, gl::sub(0, last_addr));
        return;
    }
    const uint64_t base = last_is_rw == 1 ? gl::sub(0, SPAN) : gl::add(last_addr, 1);
    const uint64_t addr = gl::add(base, gl::canon((uint64_t)(i - filled)));
    const uint64_t d = (i == filled) ? gl::sub(addr, last_addr) : 1;
    put(16, 1), put(1, last_env), put(3, addr), put(17, 1);
    put(19, d), put(20, (i == filled) ? gl::inv(d) : 1);
    put(22, gl::sub(0, addr)), put(24, 1), put(26, gl::sub(0, addr));
}
void memory_trace(ola_ctx* ctx, const uint64_t* d_cells, size_t ncells, uint32_t log_n, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    OLA_CHECK(log_n >= 1 && ncells <= n, OLA_ERR_INVALID_ARG, "Memory table: at least two rows and room for every cell");
    Launch lz(ctx, "gen_memory_fill");
    memory_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_cells, ncells, n, d_out);
    check_launch("memory_fill_kernel");
}

// ---- generate_prog_trace (circuits/src/generation/prog.rs:18-157) ----------------------------------------------------------------------
// The Program table: one row per fetched program word on the executed side (one or two per executed main line: the instruction
// and, when it carries one, its immediate -- a prefix sum places them), one row per word of every program on the other side, the
// beta-compressed (code address, pc, word) of both, and the permuted pair of the lookup between them (permuted_cols at the
// table's full size: 2^23 rows for the benchmark's run).
__device__ __forceinline__ uint64_t prog_compress(const uint64_t v[6], uint64_t beta) {  // sum_k v[k] beta^k (prog.rs:149-156)
    uint64_t acc = 0;
#pragma unroll
    for (int k = 5; k >= 0; --k) acc = gl::add(gl::mul(acc, beta), gl::canon(v[k]));
    return acc;
}
__device__ __forceinline__ bool step_fetches_imm(const uint64_t* s) { return s[26] == 1 || s[27] == (1ull << 22) || s[27] == (1ull << 21); }
__global__ void prog_count_kernel(const uint64_t* __restrict__ steps, size_t nsteps, uint32_t* __restrict__ w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsteps) return;
    const uint64_t* s = steps + i * 66;
    w[i] = s[13] == 1 ? 0u : (step_fetches_imm(s) ? 2u : 1u);  // ext lines fetch nothing (prog.rs:58-60)
}
__global__ void prog_exec_fill_kernel(const uint64_t* __restrict__ steps, size_t nsteps, const uint32_t* __restrict__ at, uint64_t beta, size_t n,
                                      uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsteps) return;
    const uint64_t* s = steps + i * 66;
    if (s[13] == 1) return;
    size_t e = at[i];
    uint64_t v[6] = {s[6], s[7], s[8], s[9], s[12], s[25]};
    for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
        for (int k = 0; k < 6; ++k) out[(size_t)(8 + k) * n + e] = gl::canon(v[k]);
        out[(size_t)14 * n + e] = prog_compress(v, beta);
        out[(size_t)16 * n + e] = 1;
        if (rep == 1 || !step_fetches_imm(s)) break;
        v[4] = s[12] + 1;  // pc + 1: the immediate word
        v[5] = s[28];
        ++e;
    }
}
__global__ void prog_rows_fill_kernel(const uint64_t* __restrict__ rows /* [m][6] */, size_t m, uint64_t beta, size_t n, uint64_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    uint64_t v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        v[k] = rows[j * 6 + k];
        out[(size_t)k * n + j] = gl::canon(v[k]);
    }
    out[(size_t)6 * n + j] = prog_compress(v, beta);
    out[(size_t)17 * n + j] = 1;
}
// roots[8] (host) = start_root[4], end_root[4]; returns beta.  d_out = [18][n]
uint64_t program_trace(ola_ctx* ctx, const uint64_t* d_steps, size_t nsteps, const uint64_t* d_prog_rows, size_t m, const uint64_t* roots, uint32_t log_n,
                       uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    OLA_CHECK(log_n >= 1 && log_n <= 24 && m <= n, OLA_ERR_INVALID_ARG, "Program table: a power of two of at least 2 rows with room for every program word");
    stark::Challenger ch(OLA_HASH_POSEIDON);  // observe start[i], end[i] for i in 0..4 (prog.rs:23-29)
    for (int i = 0; i < 4; ++i) {
        ch.observe(gl::canon(roots[i]));
        ch.observe(gl::canon(roots[4 + i]));
    }
    const uint64_t beta = ch.get_challenge();
    OLA_CUDA(cudaMemsetAsync(d_out, 0, 18 * n * sizeof(uint64_t), ctx->stream));
    if (nsteps) {
        const size_t nb = (nsteps + SCAN_BLOCK - 1) / SCAN_BLOCK;
        Buf w32((2 * nsteps + nb + 8) / 2 + 4);
        uint32_t* w = w32.u32();
        uint32_t* at = w + nsteps;
        uint32_t* sums = at + nsteps;
        uint32_t* total = sums + nb;
        Launch lz(ctx, "gen_prog_fill");
        const unsigned gs = (unsigned)((nsteps + 255) / 256);
        prog_count_kernel<<<gs, 256, 0, ctx->stream>>>(d_steps, nsteps, w);
        exclusive_scan(ctx, w, at, sums, total, nsteps);
        uint32_t exec_len = 0;
        OLA_CUDA(cudaMemcpyAsync(&exec_len, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        OLA_CHECK(exec_len <= n, OLA_ERR_INVALID_ARG, "Program table: the executed lines fetch more words than the table has rows");
        prog_exec_fill_kernel<<<gs, 256, 0, ctx->stream>>>(d_steps, nsteps, at, beta, n, d_out);
        check_launch("prog_exec_fill_kernel");
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (m) {
        Launch lz(ctx, "gen_prog_fill");
        prog_rows_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(d_prog_rows, m, beta, n, d_out);
        check_launch("prog_rows_fill_kernel");
    }
    permuted_cols(ctx, d_out + 14 * n, d_out + 6 * n, n, d_out + 15 * n, d_out + 7 * n);  // prog.rs:132-135
    return beta;
}

}  // namespace lookup
}  // namespace ola
