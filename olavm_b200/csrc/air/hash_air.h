// AIR constraints of the Poseidon, PoseidonChunk and Bitwise tables, single transcription shared by the device
// quotient kernels and the oracle (see the note at the top of cpu_air.h).  Constraints in the reference's source order.
//   Poseidon       circuits/src/builtins/poseidon/poseidon_stark.rs:61-143       (columns.rs:6-42, degree 7)
//                  with the field-generic round helpers of core/src/util/poseidon_utils.rs:289-376
//   PoseidonChunk  circuits/src/builtins/poseidon/poseidon_chunk_stark.rs:100-277 (columns.rs:44-71, degree 3)
//   Bitwise        circuits/src/builtins/bitwise/bitwise_stark.rs:44-186          (columns.rs:23-48, degree 3, compress challenge)
#pragma once
#include "builtins_air.h"

namespace ola {
namespace air {

namespace psdn {
enum : int {
    FILTER_LOOKED_NORMAL = 0, FILTER_LOOKED_TREEKEY = 1, FILTER_LOOKED_STORAGE_LEAF = 2, FILTER_LOOKED_STORAGE_BRANCH = 3,
    COL_POSEIDON_INPUT = 4, COL_POSEIDON_OUTPUT = 16, COL_POSEIDON_FULL_ROUND_0_1_STATE = 28, COL_POSEIDON_FULL_ROUND_0_2_STATE = 40,
    COL_POSEIDON_FULL_ROUND_0_3_STATE = 52, COL_POSEIDON_PARTIAL_ROUND_ELEMENT = 64, COL_POSEIDON_FULL_ROUND_1_0_STATE = 86,
    COL_POSEIDON_FULL_ROUND_1_1_STATE = 98, COL_POSEIDON_FULL_ROUND_1_2_STATE = 110, COL_POSEIDON_FULL_ROUND_1_3_STATE = 122,
    NUM_POSEIDON_COLS = 134
};
// poseidon_utils.rs:295-300
template <class T>
AIR_FN T sbox_monomial(const T x) {
    const T x2 = x * x, x4 = x2 * x2, x3 = x * x2;
    return x3 * x4;
}
// K supplies the Poseidon parameter tables as canonical u64:
//   K::round(i) ALL_ROUND_CONSTANTS[i], K::circ(i), K::diag(i), K::first(i) FAST_PARTIAL_FIRST_ROUND_CONSTANT[i],
//   K::partial(r) FAST_PARTIAL_ROUND_CONSTANTS[r], K::init(r, c) FAST_PARTIAL_ROUND_INITIAL_MATRIX[r][c],
//   K::what(r, i) FAST_PARTIAL_ROUND_W_HATS[r][i], K::vs(r, i) FAST_PARTIAL_ROUND_VS[r][i]
template <class T, class K>
AIR_FN void mds_layer_field(T* state) {  // poseidon_utils.rs:308-326
    T res[12];
    for (int r = 0; r < 12; ++r) {
        T acc = kc<T>(0);
        for (int i = 0; i < 12; ++i) acc = acc + state[(i + r) % 12] * kc<T>(K::circ(i));
        acc = acc + state[r] * kc<T>(K::diag(r));
        res[r] = acc;
    }
    for (int r = 0; r < 12; ++r) state[r] = res[r];
}
template <class T, class K>
AIR_FN void mds_partial_layer_fast_field(T* state, int r) {  // poseidon_utils.rs:358-376
    const T s0 = state[0];
    T d = s0 * kc<T>(K::circ(0) + K::diag(0));
    for (int i = 1; i < 12; ++i) d = d + state[i] * kc<T>(K::what(r, i - 1));
    for (int i = 1; i < 12; ++i) state[i] = s0 * kc<T>(K::vs(r, i - 1)) + state[i];
    state[0] = d;
}
template <class T, class R, class C, class K>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    (void)nv;
    const T one = kc<T>(1);
    for (int k = 9; k < 12; ++k) {
        const T cap = lv[COL_POSEIDON_INPUT + k];
        yc.constraint(lv[FILTER_LOOKED_TREEKEY] * cap);
        yc.constraint(lv[FILTER_LOOKED_STORAGE_LEAF] * cap);
        yc.constraint(lv[FILTER_LOOKED_STORAGE_BRANCH] * cap);
    }
    yc.constraint(lv[FILTER_LOOKED_STORAGE_LEAF] * (one - lv[COL_POSEIDON_INPUT + 8]));

    T state[12];
    for (int i = 0; i < 12; ++i) state[i] = lv[COL_POSEIDON_INPUT + i];
    int round_ctr = 0;
    // first set of full rounds
    for (int r = 0; r < 4; ++r) {
        for (int i = 0; i < 12; ++i) state[i] = state[i] + kc<T>(K::round(i + 12 * round_ctr));
        if (r != 0) {
            const int base = r == 1 ? COL_POSEIDON_FULL_ROUND_0_1_STATE : r == 2 ? COL_POSEIDON_FULL_ROUND_0_2_STATE : COL_POSEIDON_FULL_ROUND_0_3_STATE;
            for (int i = 0; i < 12; ++i) {
                const T sbox_in = lv[base + i];
                yc.constraint(state[i] - sbox_in);
                state[i] = sbox_in;
            }
        }
        for (int i = 0; i < 12; ++i) state[i] = sbox_monomial<T>(state[i]);
        mds_layer_field<T, K>(state);
        round_ctr += 1;
    }
    // partial rounds: partial_first_constant_layer, mds_partial_layer_init (poseidon_utils.rs:328-356)
    for (int i = 0; i < 12; ++i) state[i] = state[i] + kc<T>(K::first(i));
    {
        T result[12];
        result[0] = state[0];
        for (int c = 1; c < 12; ++c) result[c] = kc<T>(0);
        for (int r = 1; r < 12; ++r)
            for (int c = 1; c < 12; ++c) result[c] = result[c] + state[r] * kc<T>(K::init(r - 1, c - 1));
        for (int c = 0; c < 12; ++c) state[c] = result[c];
    }
    for (int r = 0; r < 21; ++r) {
        const T sbox_in = lv[COL_POSEIDON_PARTIAL_ROUND_ELEMENT + r];
        yc.constraint(state[0] - sbox_in);
        state[0] = sbox_monomial<T>(sbox_in);
        state[0] = state[0] + kc<T>(K::partial(r));
        mds_partial_layer_fast_field<T, K>(state, r);
    }
    {
        const T sbox_in = lv[COL_POSEIDON_PARTIAL_ROUND_ELEMENT + 21];
        yc.constraint(state[0] - sbox_in);
        state[0] = sbox_monomial<T>(sbox_in);
        mds_partial_layer_fast_field<T, K>(state, 21);
    }
    round_ctr += 22;
    // second set of full rounds
    for (int r = 0; r < 4; ++r) {
        for (int i = 0; i < 12; ++i) state[i] = state[i] + kc<T>(K::round(i + 12 * round_ctr));
        const int base = COL_POSEIDON_FULL_ROUND_1_0_STATE + 12 * r;
        for (int i = 0; i < 12; ++i) {
            const T sbox_in = lv[base + i];
            yc.constraint(state[i] - sbox_in);
            state[i] = sbox_in;
        }
        for (int i = 0; i < 12; ++i) state[i] = sbox_monomial<T>(state[i]);
        mds_layer_field<T, K>(state);
        round_ctr += 1;
    }
    for (int i = 0; i < 12; ++i) yc.constraint(state[i] - lv[COL_POSEIDON_OUTPUT + i]);
}
}  // namespace psdn

namespace psdn_chunk {
enum : int {
    COL_POSEIDON_CHUNK_TX_IDX = 0, COL_POSEIDON_CHUNK_ENV_IDX = 1, COL_POSEIDON_CHUNK_CLK = 2, COL_POSEIDON_CHUNK_OPCODE = 3, COL_POSEIDON_CHUNK_OP0 = 4,
    COL_POSEIDON_CHUNK_OP1 = 5, COL_POSEIDON_CHUNK_DST = 6, COL_POSEIDON_CHUNK_ACC_CNT = 7, COL_POSEIDON_CHUNK_VALUE = 8, COL_POSEIDON_CHUNK_CAP = 16,
    COL_POSEIDON_CHUNK_HASH = 20, COL_POSEIDON_CHUNK_IS_EXT_LINE = 32, COL_POSEIDON_CHUNK_IS_RESULT_LINE = 33, COL_POSEIDON_CHUNK_IS_FIRST_PADDING = 34,
    COL_POSEIDON_CHUNK_FILTER_LOOKED_CPU = 42, COL_POSEIDON_CHUNK_FILTER_LOOKING_MEM = 43, COL_POSEIDON_CHUNK_FILTER_LOOKING_POSEIDON = 51,
    COL_POSEIDON_CHUNK_IS_PADDING_LINE = 52, NUM_POSEIDON_CHUNK_COLS = 53
};
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    const T one = kc<T>(1);
    const T lv_pad = lv[COL_POSEIDON_CHUNK_IS_PADDING_LINE], nv_pad = nv[COL_POSEIDON_CHUNK_IS_PADDING_LINE];
    const T lv_ext = lv[COL_POSEIDON_CHUNK_IS_EXT_LINE], nv_ext = nv[COL_POSEIDON_CHUNK_IS_EXT_LINE];
    yc.constraint(lv_pad * (one - lv_pad));
    yc.constraint_transition((nv_pad - lv_pad) * (nv_pad - lv_pad - one));
    yc.constraint(lv_ext * (one - lv_ext));
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_TX_IDX] - lv[COL_POSEIDON_CHUNK_TX_IDX]));
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_ENV_IDX] - lv[COL_POSEIDON_CHUNK_ENV_IDX]));
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_CLK] - lv[COL_POSEIDON_CHUNK_CLK]));
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_OPCODE] - lv[COL_POSEIDON_CHUNK_OPCODE]));
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_OP1] - lv[COL_POSEIDON_CHUNK_OP1]));
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_DST] - lv[COL_POSEIDON_CHUNK_DST]));
    yc.constraint_first_row((one - lv_pad) * lv_ext);
    T sum_is_first_padding = kc<T>(0);
    for (int i = 0; i < 8; ++i) {
        const T v = lv[COL_POSEIDON_CHUNK_IS_FIRST_PADDING + i];
        yc.constraint(v * (one - v));
        sum_is_first_padding = sum_is_first_padding + v;
    }
    yc.constraint(sum_is_first_padding * (one - sum_is_first_padding));
    T v_line_acc_addends[8];
    T n_v_line_acc_total_addend = kc<T>(0);
    {
        T s = kc<T>(0), ns = kc<T>(0);
        for (int i = 0; i < 8; ++i) {
            s = s + lv[COL_POSEIDON_CHUNK_IS_FIRST_PADDING + i];
            ns = ns + nv[COL_POSEIDON_CHUNK_IS_FIRST_PADDING + i];
            v_line_acc_addends[i] = one - s;
            n_v_line_acc_total_addend = n_v_line_acc_total_addend + (one - ns);
        }
    }
    yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_ACC_CNT] - lv[COL_POSEIDON_CHUNK_ACC_CNT] - n_v_line_acc_total_addend));
    yc.constraint(sum_is_first_padding * nv_ext);
    yc.constraint(sum_is_first_padding * (one - lv[COL_POSEIDON_CHUNK_IS_RESULT_LINE]));
    yc.constraint(sum_is_first_padding * (lv[COL_POSEIDON_CHUNK_ACC_CNT] - lv[COL_POSEIDON_CHUNK_OP1]));
    yc.constraint((lv[COL_POSEIDON_CHUNK_ACC_CNT] - lv[COL_POSEIDON_CHUNK_OP1]) * (one - nv_ext));
    for (int i = 0; i < 12; ++i) yc.constraint((one - lv_ext) * lv[COL_POSEIDON_CHUNK_HASH + i]);
    for (int i = 0; i < 4; ++i) yc.constraint(nv_ext * (nv[COL_POSEIDON_CHUNK_CAP + i] - lv[COL_POSEIDON_CHUNK_HASH + 8 + i]));
    yc.constraint((one - lv_ext) * nv_ext * (nv[COL_POSEIDON_CHUNK_OP0] - lv[COL_POSEIDON_CHUNK_OP0]));
    yc.constraint(lv_ext * nv_ext * (nv[COL_POSEIDON_CHUNK_OP0] - lv[COL_POSEIDON_CHUNK_OP0] - kc<T>(8)));
    yc.constraint((one - lv_pad) * (one - lv_ext) * (one - lv[COL_POSEIDON_CHUNK_FILTER_LOOKED_CPU]));
    yc.constraint((one - lv_pad) * lv_ext * lv[COL_POSEIDON_CHUNK_FILTER_LOOKED_CPU]);
    yc.constraint(lv_pad * lv[COL_POSEIDON_CHUNK_FILTER_LOOKED_CPU]);
    for (int i = 0; i < 8; ++i) {
        const T filter = lv[COL_POSEIDON_CHUNK_FILTER_LOOKING_MEM + i];
        yc.constraint((one - lv_ext) * filter);
        yc.constraint(lv_ext * (filter - v_line_acc_addends[i]));
    }
    yc.constraint((one - lv_pad) * lv_ext * (one - lv[COL_POSEIDON_CHUNK_FILTER_LOOKING_POSEIDON]));
    yc.constraint((one - lv_pad) * (one - lv_ext) * lv[COL_POSEIDON_CHUNK_FILTER_LOOKING_POSEIDON]);
}
}  // namespace psdn_chunk

namespace bitwise {
enum : int {
    FILTER = 0, TAG = 1, OP0 = 2, OP1 = 3, RES = 4, OP0_LIMBS = 5, OP1_LIMBS = 9, RES_LIMBS = 13, OP0_LIMBS_PERMUTED = 17, OP1_LIMBS_PERMUTED = 21,
    RES_LIMBS_PERMUTED = 25, COMPRESS_LIMBS = 29, COMPRESS_PERMUTED = 33, FIX_RANGE_CHECK_U8 = 37, FIX_RANGE_CHECK_U8_PERMUTED = 38, FIX_TAG = 50,
    FIX_BITWSIE_OP0 = 51, FIX_BITWSIE_OP1 = 52, FIX_BITWSIE_RES = 53, FIX_COMPRESS = 54, FIX_COMPRESS_PERMUTED = 55, COL_NUM_BITWISE = 59
};
// reduce_with_powers (plonk/plonk_common.rs): sum_i terms[i] * alpha^i
template <class T, class R>
AIR_FN T reduce_with_powers4(const R& lv, int start, const T alpha) {
    T sum = kc<T>(0);
    for (int i = 3; i >= 0; --i) sum = sum * alpha + lv[start + i];
    return sum;
}
// beta = the table's compress challenge (BitwiseStark::get_compress_challenge, bitwise_stark.rs:36-38)
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc, const T beta) {
    const T base = kc<T>(1 << 8);
    yc.constraint(reduce_with_powers4<T, R>(lv, OP0_LIMBS, base) - lv[OP0]);
    yc.constraint(reduce_with_powers4<T, R>(lv, OP1_LIMBS, base) - lv[OP1]);
    yc.constraint(reduce_with_powers4<T, R>(lv, RES_LIMBS, base) - lv[RES]);
    for (int i = 0; i < 4; ++i)
        yc.constraint(lv[TAG] + lv[OP0_LIMBS + i] * beta + lv[OP1_LIMBS + i] * beta * beta + lv[RES_LIMBS + i] * beta * beta * beta - lv[COMPRESS_LIMBS + i]);
    for (int i = 0; i < 4; ++i) eval_lookups_t<T, R, C>(lv, nv, yc, OP0_LIMBS_PERMUTED + i, FIX_RANGE_CHECK_U8_PERMUTED + i);
    for (int i = 0; i < 4; ++i) eval_lookups_t<T, R, C>(lv, nv, yc, OP1_LIMBS_PERMUTED + i, FIX_RANGE_CHECK_U8_PERMUTED + 4 + i);
    for (int i = 0; i < 4; ++i) eval_lookups_t<T, R, C>(lv, nv, yc, RES_LIMBS_PERMUTED + i, FIX_RANGE_CHECK_U8_PERMUTED + 8 + i);
    for (int i = 0; i < 4; ++i) eval_lookups_t<T, R, C>(lv, nv, yc, COMPRESS_PERMUTED + i, FIX_COMPRESS_PERMUTED + i);
}
}  // namespace bitwise

}  // namespace air
}  // namespace ola
