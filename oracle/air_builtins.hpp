/* ORACLE (test infrastructure, NOT product code) -- AIRs of the builtin tables, restated from the Rust on their own
 * (nothing under olavm_b200/ is included).  Constraints in source order.
 *   Bitwise        circuits/src/builtins/bitwise/{columns.rs:23-71, bitwise_stark.rs:44-186}; reduce_with_powers
 *                  plonky2/plonky2/src/plonk/plonk_common.rs:116-128
 *   Tape           circuits/src/builtins/tape/{columns.rs:3-9, tape_stark.rs:48-137}
 *   SCCall         circuits/src/builtins/sccall/{columns.rs, sccall_stark.rs:60-75}
 *   Poseidon       circuits/src/builtins/poseidon/{columns.rs:6-42, poseidon_stark.rs:61-143} with the field-generic round
 *                  helpers of core/src/util/poseidon_utils.rs:289-376; parameter tables oracle/poseidon_constants.h
 *   PoseidonChunk  circuits/src/builtins/poseidon/{columns.rs:44-71, poseidon_chunk_stark.rs:100-277} */
#ifndef ORC_AIR_BUILTINS_HPP
#define ORC_AIR_BUILTINS_HPP
#include "poseidon_constants.h"
#include "stark.hpp"

namespace orc {

/* circuits/src/stark/lookup.rs:13-35 */
template <class O>
void air_eval_lookups(const P<O>* lv, const P<O>* nv, Consumer<O>& yc, int col_permuted_input, int col_permuted_table) {
    (void)lv;
    const P<O> local_perm_input = lv[col_permuted_input];
    const P<O> next_perm_table = nv[col_permuted_table];
    const P<O> next_perm_input = nv[col_permuted_input];
    const P<O> diff_input_prev = next_perm_input - local_perm_input;
    const P<O> diff_input_table = next_perm_input - next_perm_table;
    yc.constraint(diff_input_prev * diff_input_table);
    yc.constraint_last_row(diff_input_table);
}

namespace bitwise_air {
enum {
    FILTER = 0, TAG, OP0, OP1, RES,
    OP0_LIMBS = RES + 1, OP1_LIMBS = OP0_LIMBS + 4, RES_LIMBS = OP1_LIMBS + 4,
    OP0_LIMBS_PERMUTED = RES_LIMBS + 4, OP1_LIMBS_PERMUTED = OP0_LIMBS_PERMUTED + 4, RES_LIMBS_PERMUTED = OP1_LIMBS_PERMUTED + 4,
    COMPRESS_LIMBS = RES_LIMBS_PERMUTED + 4, COMPRESS_PERMUTED = COMPRESS_LIMBS + 4,
    FIX_RANGE_CHECK_U8 = COMPRESS_PERMUTED + 4, FIX_RANGE_CHECK_U8_PERMUTED = FIX_RANGE_CHECK_U8 + 1, /* 12 columns */
    FIX_TAG = FIX_RANGE_CHECK_U8_PERMUTED + 12, FIX_BITWSIE_OP0, FIX_BITWSIE_OP1, FIX_BITWSIE_RES, FIX_COMPRESS,
    FIX_COMPRESS_PERMUTED = FIX_COMPRESS + 1, /* 4 columns */
    NUM_COLS = FIX_COMPRESS_PERMUTED + 4
};
static_assert(NUM_COLS == 59, "bitwise/columns.rs layout");

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc, P<O> beta) {
    typedef P<O> T;
    const T base = T::c(1 << 8);
    const int limbs[3] = {OP0_LIMBS, OP1_LIMBS, RES_LIMBS}, whole[3] = {OP0, OP1, RES};
    for (int g = 0; g < 3; g++) {
        T sum = T::zero(); /* reduce_with_powers: Horner from the last limb */
        for (int i = 3; i >= 0; i--) sum = sum * base + lv[limbs[g] + i];
        yc.constraint(sum - lv[whole[g]]);
    }
    for (int i = 0; i < 4; i++)
        yc.constraint(lv[TAG] + lv[OP0_LIMBS + i] * beta + lv[OP1_LIMBS + i] * beta * beta + lv[RES_LIMBS + i] * beta * beta * beta - lv[COMPRESS_LIMBS + i]);
    for (int i = 0; i < 4; i++) air_eval_lookups<O>(lv, nv, yc, OP0_LIMBS_PERMUTED + i, FIX_RANGE_CHECK_U8_PERMUTED + i);
    for (int i = 0; i < 4; i++) air_eval_lookups<O>(lv, nv, yc, OP1_LIMBS_PERMUTED + i, FIX_RANGE_CHECK_U8_PERMUTED + 4 + i);
    for (int i = 0; i < 4; i++) air_eval_lookups<O>(lv, nv, yc, RES_LIMBS_PERMUTED + i, FIX_RANGE_CHECK_U8_PERMUTED + 8 + i);
    for (int i = 0; i < 4; i++) air_eval_lookups<O>(lv, nv, yc, COMPRESS_PERMUTED + i, FIX_COMPRESS_PERMUTED + i);
}
}  // namespace bitwise_air

namespace tape_air {
enum { TX_IDX = 0, IS_INIT_SEG, OPCODE, ADDR, VALUE, FILTER_LOOKED, NUM_COLS };
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    const T ONE = T::one();
    const T op_tload = T::c((uint64_t)1 << 9), op_tstore = T::c((uint64_t)1 << 8), op_sccall = T::c((uint64_t)1 << 7);
    yc.constraint(lv[OPCODE] * (lv[OPCODE] - op_tstore) * (lv[OPCODE] - op_tload) * (lv[OPCODE] - op_sccall));
    yc.constraint_first_row(lv[TX_IDX]);
    yc.constraint_transition((nv[TX_IDX] - lv[TX_IDX]) * (nv[TX_IDX] - lv[TX_IDX] - ONE));
    const T is_in_same_tx = ONE - (nv[TX_IDX] - lv[TX_IDX]);
    yc.constraint(lv[IS_INIT_SEG] * (ONE - lv[IS_INIT_SEG]));
    yc.constraint_transition((ONE - is_in_same_tx) * (ONE - nv[IS_INIT_SEG]));
    yc.constraint_transition(is_in_same_tx * (nv[IS_INIT_SEG] - lv[IS_INIT_SEG]) * (lv[IS_INIT_SEG] - nv[IS_INIT_SEG] - ONE));
    yc.constraint(lv[IS_INIT_SEG] * lv[OPCODE] * (lv[OPCODE] - op_tload));
    yc.constraint((ONE - lv[IS_INIT_SEG]) * (lv[OPCODE] - op_tload) * (lv[OPCODE] - op_tstore) * (lv[OPCODE] - op_sccall));
    yc.constraint_first_row(lv[ADDR]);
    yc.constraint_transition((ONE - is_in_same_tx) * nv[ADDR]);
    yc.constraint_transition(is_in_same_tx * (nv[ADDR] - lv[ADDR]) * (nv[ADDR] - lv[ADDR] - ONE));
    yc.constraint_transition(is_in_same_tx * (ONE - (nv[ADDR] - lv[ADDR])) * (nv[VALUE] - lv[VALUE]));
    yc.constraint_transition(is_in_same_tx * (ONE - (nv[ADDR] - lv[ADDR])) * (nv[OPCODE] - op_tload));
    yc.constraint(is_in_same_tx * (nv[ADDR] - lv[ADDR]) * nv[OPCODE] * (nv[OPCODE] - op_tstore) * (nv[OPCODE] - op_sccall));
    yc.constraint(lv[OPCODE] * (lv[OPCODE] - op_tload) * (ONE - lv[FILTER_LOOKED]));
}
}  // namespace tape_air

namespace sccall_air {
enum {
    TX_IDX = 0, CALLER_ENV_IDX,
    CALLER_EXE_CTX = CALLER_ENV_IDX + 1, /* 4 */
    CALLER_CODE_CTX = CALLER_EXE_CTX + 4, /* 4 */
    CALLER_OP1_IMM = CALLER_CODE_CTX + 4, CLK_CALLER_CALL, CLK_CALLER_RET,
    CALLER_REG = CLK_CALLER_RET + 1, /* 10 */
    CALLEE_ENV_IDX = CALLER_REG + 10, CLK_CALLEE_END, IS_PADDING, NUM_COLS
};
static_assert(NUM_COLS == 26, "sccall/columns.rs layout");
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    (void)nv;
    yc.constraint(lv[CLK_CALLER_RET] - lv[CLK_CALLER_CALL] - lv[CALLER_OP1_IMM]);
}
}  // namespace sccall_air

namespace poseidon_air {
enum {
    FILTER_LOOKED_NORMAL = 0, FILTER_LOOKED_TREEKEY, FILTER_LOOKED_STORAGE_LEAF, FILTER_LOOKED_STORAGE_BRANCH,
    INPUT = FILTER_LOOKED_STORAGE_BRANCH + 1, OUTPUT = INPUT + 12,
    FULL_0_1 = OUTPUT + 12, FULL_0_2 = FULL_0_1 + 12, FULL_0_3 = FULL_0_2 + 12,
    PARTIAL = FULL_0_3 + 12, /* 22 */
    FULL_1_0 = PARTIAL + 22, FULL_1_1 = FULL_1_0 + 12, FULL_1_2 = FULL_1_1 + 12, FULL_1_3 = FULL_1_2 + 12,
    NUM_COLS = FULL_1_3 + 12
};
static_assert(NUM_COLS == 134, "poseidon/columns.rs layout");
enum { WIDTH = 12, HALF_N_FULL_ROUNDS = 4, N_PARTIAL_ROUNDS = 22 };

template <class T>
T sbox_monomial(T x) {
    const T x2 = x * x, x4 = x2 * x2, x3 = x * x2;
    return x3 * x4;
}
template <class T>
void constant_layer_field(T* state, int round_ctr) {
    for (int i = 0; i < 12; i++) state[i] = state[i] + T::c(ORC_ALL_ROUND_CONSTANTS[i + 12 * round_ctr]);
}
template <class T>
void mds_layer_field(T* state) {
    T res[WIDTH];
    for (int r = 0; r < WIDTH; r++) {
        T acc = T::zero();
        for (int i = 0; i < WIDTH; i++) acc = acc + state[(i + r) % WIDTH] * T::c(ORC_MDS_MATRIX_CIRC[i]);
        acc = acc + state[r] * T::c(ORC_MDS_MATRIX_DIAG[r]);
        res[r] = acc;
    }
    for (int r = 0; r < WIDTH; r++) state[r] = res[r];
}
template <class T>
void mds_partial_layer_init(T* state) {
    T result[WIDTH];
    for (int c = 0; c < WIDTH; c++) result[c] = T::zero();
    result[0] = state[0];
    for (int r = 1; r < WIDTH; r++)
        for (int c = 1; c < WIDTH; c++) result[c] = result[c] + state[r] * T::c(ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r - 1][c - 1]);
    for (int c = 0; c < WIDTH; c++) state[c] = result[c];
}
template <class T>
void mds_partial_layer_fast_field(T* state, int r) {
    const T s0 = state[0];
    T d = s0 * T::c(ORC_MDS_MATRIX_CIRC[0] + ORC_MDS_MATRIX_DIAG[0]);
    for (int i = 1; i < WIDTH; i++) d = d + state[i] * T::c(ORC_FAST_PARTIAL_ROUND_W_HATS[r][i - 1]);
    T result[WIDTH];
    result[0] = d;
    for (int i = 1; i < WIDTH; i++) result[i] = state[0] * T::c(ORC_FAST_PARTIAL_ROUND_VS[r][i - 1]) + state[i];
    for (int i = 0; i < WIDTH; i++) state[i] = result[i];
}

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    (void)nv;
    typedef P<O> T;
    for (int k = 9; k < 12; k++) {
        const T cap = lv[INPUT + k];
        yc.constraint(lv[FILTER_LOOKED_TREEKEY] * cap);
        yc.constraint(lv[FILTER_LOOKED_STORAGE_LEAF] * cap);
        yc.constraint(lv[FILTER_LOOKED_STORAGE_BRANCH] * cap);
    }
    yc.constraint(lv[FILTER_LOOKED_STORAGE_LEAF] * (T::one() - lv[INPUT + 8]));

    T state[WIDTH];
    for (int i = 0; i < WIDTH; i++) state[i] = lv[INPUT + i];
    int round_ctr = 0;
    const int full0[4] = {-1, FULL_0_1, FULL_0_2, FULL_0_3}, full1[4] = {FULL_1_0, FULL_1_1, FULL_1_2, FULL_1_3};
    for (int r = 0; r < HALF_N_FULL_ROUNDS; r++) {
        constant_layer_field(state, round_ctr);
        if (r != 0)
            for (int i = 0; i < WIDTH; i++) {
                const T sbox_in = lv[full0[r] + i];
                yc.constraint(state[i] - sbox_in);
                state[i] = sbox_in;
            }
        for (int i = 0; i < WIDTH; i++) state[i] = sbox_monomial(state[i]);
        mds_layer_field(state);
        round_ctr += 1;
    }
    for (int i = 0; i < WIDTH; i++) state[i] = state[i] + T::c(ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
    mds_partial_layer_init(state);
    for (int r = 0; r < N_PARTIAL_ROUNDS - 1; r++) {
        const T sbox_in = lv[PARTIAL + r];
        yc.constraint(state[0] - sbox_in);
        state[0] = sbox_monomial(sbox_in);
        state[0] = state[0] + T::c(ORC_FAST_PARTIAL_ROUND_CONSTANTS[r]);
        mds_partial_layer_fast_field(state, r);
    }
    {
        const T sbox_in = lv[PARTIAL + N_PARTIAL_ROUNDS - 1];
        yc.constraint(state[0] - sbox_in);
        state[0] = sbox_monomial(sbox_in);
        mds_partial_layer_fast_field(state, N_PARTIAL_ROUNDS - 1);
    }
    round_ctr += N_PARTIAL_ROUNDS;
    for (int r = 0; r < HALF_N_FULL_ROUNDS; r++) {
        constant_layer_field(state, round_ctr);
        for (int i = 0; i < WIDTH; i++) {
            const T sbox_in = lv[full1[r] + i];
            yc.constraint(state[i] - sbox_in);
            state[i] = sbox_in;
        }
        for (int i = 0; i < WIDTH; i++) state[i] = sbox_monomial(state[i]);
        mds_layer_field(state);
        round_ctr += 1;
    }
    for (int i = 0; i < WIDTH; i++) yc.constraint(state[i] - lv[OUTPUT + i]);
}
}  // namespace poseidon_air

namespace poseidon_chunk_air {
enum {
    TX_IDX = 0, ENV_IDX, CLK, OPCODE, OP0, OP1, DST, ACC_CNT,
    VALUE = ACC_CNT + 1, /* 8 */
    CAP = VALUE + 8,     /* 4 */
    HASH = CAP + 4,      /* 12 */
    IS_EXT_LINE = HASH + 12, IS_RESULT_LINE,
    IS_FIRST_PADDING = IS_RESULT_LINE + 1, /* 8 */
    FILTER_LOOKED_CPU = IS_FIRST_PADDING + 8,
    FILTER_LOOKING_MEM = FILTER_LOOKED_CPU + 1, /* 8 */
    FILTER_LOOKING_POSEIDON = FILTER_LOOKING_MEM + 8, IS_PADDING_LINE, NUM_COLS
};
static_assert(NUM_COLS == 53, "poseidon/columns.rs chunk layout");

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    const T ONE = T::one();
    yc.constraint(lv[IS_PADDING_LINE] * (ONE - lv[IS_PADDING_LINE]));
    yc.constraint_transition((nv[IS_PADDING_LINE] - lv[IS_PADDING_LINE]) * (nv[IS_PADDING_LINE] - lv[IS_PADDING_LINE] - ONE));
    yc.constraint(lv[IS_EXT_LINE] * (ONE - lv[IS_EXT_LINE]));
    const int same[6] = {TX_IDX, ENV_IDX, CLK, OPCODE, OP1, DST};
    for (int i = 0; i < 6; i++) yc.constraint(nv[IS_EXT_LINE] * (nv[same[i]] - lv[same[i]]));
    yc.constraint_first_row((ONE - lv[IS_PADDING_LINE]) * lv[IS_EXT_LINE]);
    for (int i = 0; i < 8; i++) yc.constraint(lv[IS_FIRST_PADDING + i] * (ONE - lv[IS_FIRST_PADDING + i]));
    T sum_is_first_padding = T::zero();
    for (int i = 0; i < 8; i++) sum_is_first_padding = sum_is_first_padding + lv[IS_FIRST_PADDING + i];
    yc.constraint(sum_is_first_padding * (ONE - sum_is_first_padding));

    /* 1 - running sum of the first-padding flags: position i still carries a value */
    T v_line_acc_addends[8], n_v_line_acc_addends[8];
    {
        T run = T::zero(), nrun = T::zero();
        for (int i = 0; i < 8; i++) {
            run = run + lv[IS_FIRST_PADDING + i];
            nrun = nrun + nv[IS_FIRST_PADDING + i];
            v_line_acc_addends[i] = ONE - run;
            n_v_line_acc_addends[i] = ONE - nrun;
        }
    }
    T n_v_line_acc_total_addend = T::zero();
    for (int i = 0; i < 8; i++) n_v_line_acc_total_addend = n_v_line_acc_total_addend + n_v_line_acc_addends[i];
    yc.constraint(nv[IS_EXT_LINE] * (nv[ACC_CNT] - lv[ACC_CNT] - n_v_line_acc_total_addend));
    yc.constraint(sum_is_first_padding * nv[IS_EXT_LINE]);
    yc.constraint(sum_is_first_padding * (ONE - lv[IS_RESULT_LINE]));
    yc.constraint(sum_is_first_padding * (lv[ACC_CNT] - lv[OP1]));
    yc.constraint((lv[ACC_CNT] - lv[OP1]) * (ONE - nv[IS_EXT_LINE]));
    for (int i = 0; i < 12; i++) yc.constraint((ONE - lv[IS_EXT_LINE]) * lv[HASH + i]);
    for (int i = 0; i < 4; i++) yc.constraint(nv[IS_EXT_LINE] * (nv[CAP + i] - lv[HASH + 8 + i]));
    yc.constraint((ONE - lv[IS_EXT_LINE]) * nv[IS_EXT_LINE] * (nv[OP0] - lv[OP0]));
    yc.constraint(lv[IS_EXT_LINE] * nv[IS_EXT_LINE] * (nv[OP0] - lv[OP0] - T::c(8)));
    yc.constraint((ONE - lv[IS_PADDING_LINE]) * (ONE - lv[IS_EXT_LINE]) * (ONE - lv[FILTER_LOOKED_CPU]));
    yc.constraint((ONE - lv[IS_PADDING_LINE]) * lv[IS_EXT_LINE] * lv[FILTER_LOOKED_CPU]);
    yc.constraint(lv[IS_PADDING_LINE] * lv[FILTER_LOOKED_CPU]);
    for (int i = 0; i < 8; i++) {
        const T filter = lv[FILTER_LOOKING_MEM + i];
        yc.constraint((ONE - lv[IS_EXT_LINE]) * filter);
        yc.constraint(lv[IS_EXT_LINE] * (filter - v_line_acc_addends[i]));
    }
    yc.constraint((ONE - lv[IS_PADDING_LINE]) * lv[IS_EXT_LINE] * (ONE - lv[FILTER_LOOKING_POSEIDON]));
    yc.constraint((ONE - lv[IS_PADDING_LINE]) * (ONE - lv[IS_EXT_LINE]) * lv[FILTER_LOOKING_POSEIDON]);
}
}  // namespace poseidon_chunk_air

}  // namespace orc
#endif
