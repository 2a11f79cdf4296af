// Trace-generation tail: the pieces of circuits/src/generation that sit directly in
// front of prove_with_traces and are worth running next to the prover.
//
//   generate_poseidon_trace  circuits/src/generation/poseidon.rs:5-130 + the round states the executor records for every
//                            hash (core/src/util/poseidon_utils.rs:289-420, PoseidonRow): one thread per table row
//                            replays the permutation in the AIR's formulation (builtins/poseidon/poseidon_stark.rs:83-141)
//                            and stores input, the S-box inputs of full rounds 1-3 / 0-3, state[0] before each of the 22
//                            partial-round S-boxes, and the output; rows past `nrows` are the zero-input row
//                            (POSEIDON_ZERO_HASH_*, generation/poseidon.rs:83-126)
//   compress challenge       the Fiat-Shamir beta of the Bitwise / Program tables: a Challenger that observes whole
//                            columns (generation/builtin.rs:118-131; generation/prog.rs:23-29).  A duplex sponge is
//                            sequential by construction, so this one is host code on the library's own transcript.
#include "air/registry.cuh"
#include "common.h"
#include "stark_types.h"

namespace ola {
namespace generation {

using air::Fp;
using K = air::PoseidonParams;

__global__ void __launch_bounds__(128) poseidon_rows_kernel(const uint64_t* __restrict__ inputs /* [nrows][12] */,
                                                            const uint64_t* __restrict__ filters /* [nrows][4] or null */, size_t nrows,
                                                            size_t n, uint64_t* __restrict__ out /* [134][n] */) {
    using namespace air::psdn;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    auto put = [&](int col, Fp v) { out[(size_t)col * n + row] = v.v; };
    Fp state[12];
    for (int i = 0; i < 12; ++i) state[i] = Fp(row < nrows ? gl::canon(inputs[row * 12 + i]) : 0);
    for (int i = 0; i < 4; ++i) put(i, Fp(row < nrows && filters ? gl::canon(filters[row * 4 + i]) : 0));
    for (int i = 0; i < 12; ++i) put(COL_POSEIDON_INPUT + i, state[i]);
    int round_ctr = 0;
    for (int r = 0; r < 4; ++r) {
        for (int i = 0; i < 12; ++i) state[i] = state[i] + Fp(K::round(i + 12 * round_ctr));
        if (r != 0)
            for (int i = 0; i < 12; ++i) put(COL_POSEIDON_FULL_ROUND_0_1_STATE + 12 * (r - 1) + i, state[i]);
        for (int i = 0; i < 12; ++i) state[i] = sbox_monomial<Fp>(state[i]);
        mds_layer_field<Fp, K>(state);
        round_ctr += 1;
    }
    for (int i = 0; i < 12; ++i) state[i] = state[i] + Fp(K::first(i));
    {
        Fp result[12];
        result[0] = state[0];
        for (int c = 1; c < 12; ++c) result[c] = Fp(0);
        for (int r = 1; r < 12; ++r)
            for (int c = 1; c < 12; ++c) result[c] = result[c] + state[r] * Fp(K::init(r - 1, c - 1));
        for (int c = 0; c < 12; ++c) state[c] = result[c];
    }
    for (int r = 0; r < 22; ++r) {
        put(COL_POSEIDON_PARTIAL_ROUND_ELEMENT + r, state[0]);
        state[0] = sbox_monomial<Fp>(state[0]);
        if (r < 21) state[0] = state[0] + Fp(K::partial(r));
        mds_partial_layer_fast_field<Fp, K>(state, r);
    }
    round_ctr += 22;
    for (int r = 0; r < 4; ++r) {
        for (int i = 0; i < 12; ++i) state[i] = state[i] + Fp(K::round(i + 12 * round_ctr));
        for (int i = 0; i < 12; ++i) put(COL_POSEIDON_FULL_ROUND_1_0_STATE + 12 * r + i, state[i]);
        for (int i = 0; i < 12; ++i) state[i] = sbox_monomial<Fp>(state[i]);
        mds_layer_field<Fp, K>(state);
        round_ctr += 1;
    }
    for (int i = 0; i < 12; ++i) put(COL_POSEIDON_OUTPUT + i, state[i]);
}

void poseidon_trace(ola_ctx* ctx, const uint64_t* d_inputs, const uint64_t* d_filters, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    const size_t n = (size_t)1 << log_n;
    Launch lz(ctx, "gen_poseidon_rows");
    poseidon_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_inputs, d_filters, nrows, n, d_out);
    check_launch("poseidon_rows_kernel");
}

}  // namespace generation
}  // namespace ola
