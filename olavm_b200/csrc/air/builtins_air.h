// AIR constraints of the builtin / program tables, single transcription shared by the device quotient kernels and
// the oracle (see the note at the top of cpu_air.h).  Constraints in the reference's source order.
//   Tape           circuits/src/builtins/tape/tape_stark.rs:45-137          (columns.rs:3-9, degree 5)
//   SCCall         circuits/src/builtins/sccall/sccall_stark.rs:72-86       (columns.rs:4-18, degree 1)
//   Program        circuits/src/program/program_stark.rs:61-100             (columns.rs:3-16, degree 3, compress challenge)
//   ProgChunk      circuits/src/program/prog_chunk_stark.rs:67-162          (columns.rs:47-62, degree 4)
//   StorageAccess  circuits/src/builtins/storage/storage_access_stark.rs:120-320 (columns.rs:3-33, degree 4)
#pragma once
#include "cpu_air.h"

namespace ola {
namespace air {

namespace tape {
enum : int { COL_TAPE_TX_IDX = 0, COL_TAPE_IS_INIT_SEG, COL_TAPE_OPCODE, COL_TAPE_ADDR, COL_TAPE_VALUE, COL_FILTER_LOOKED, NUM_COL_TAPE };
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    const T one = kc<T>(1);
    const T op_tload = kc<T>(cpu::OP_TLOAD), op_tstore = kc<T>(cpu::OP_TSTORE), op_sccall = kc<T>(cpu::OP_SCCALL);
    yc.constraint(lv[COL_TAPE_OPCODE] * (lv[COL_TAPE_OPCODE] - op_tstore) * (lv[COL_TAPE_OPCODE] - op_tload) * (lv[COL_TAPE_OPCODE] - op_sccall));
    yc.constraint_first_row(lv[COL_TAPE_TX_IDX]);
    yc.constraint_transition((nv[COL_TAPE_TX_IDX] - lv[COL_TAPE_TX_IDX]) * (nv[COL_TAPE_TX_IDX] - lv[COL_TAPE_TX_IDX] - one));
    const T is_in_same_tx = one - (nv[COL_TAPE_TX_IDX] - lv[COL_TAPE_TX_IDX]);
    yc.constraint(lv[COL_TAPE_IS_INIT_SEG] * (one - lv[COL_TAPE_IS_INIT_SEG]));
    yc.constraint_transition((one - is_in_same_tx) * (one - nv[COL_TAPE_IS_INIT_SEG]));
    yc.constraint_transition(is_in_same_tx * (nv[COL_TAPE_IS_INIT_SEG] - lv[COL_TAPE_IS_INIT_SEG]) * (lv[COL_TAPE_IS_INIT_SEG] - nv[COL_TAPE_IS_INIT_SEG] - one));
    yc.constraint(lv[COL_TAPE_IS_INIT_SEG] * lv[COL_TAPE_OPCODE] * (lv[COL_TAPE_OPCODE] - op_tload));
    yc.constraint((one - lv[COL_TAPE_IS_INIT_SEG]) * (lv[COL_TAPE_OPCODE] - op_tload) * (lv[COL_TAPE_OPCODE] - op_tstore) * (lv[COL_TAPE_OPCODE] - op_sccall));
    yc.constraint_first_row(lv[COL_TAPE_ADDR]);
    yc.constraint_transition((one - is_in_same_tx) * nv[COL_TAPE_ADDR]);
    yc.constraint_transition(is_in_same_tx * (nv[COL_TAPE_ADDR] - lv[COL_TAPE_ADDR]) * (nv[COL_TAPE_ADDR] - lv[COL_TAPE_ADDR] - one));
    yc.constraint_transition(is_in_same_tx * (one - (nv[COL_TAPE_ADDR] - lv[COL_TAPE_ADDR])) * (nv[COL_TAPE_VALUE] - lv[COL_TAPE_VALUE]));
    yc.constraint_transition(is_in_same_tx * (one - (nv[COL_TAPE_ADDR] - lv[COL_TAPE_ADDR])) * (nv[COL_TAPE_OPCODE] - op_tload));
    yc.constraint(is_in_same_tx * (nv[COL_TAPE_ADDR] - lv[COL_TAPE_ADDR]) * nv[COL_TAPE_OPCODE] * (nv[COL_TAPE_OPCODE] - op_tstore) * (nv[COL_TAPE_OPCODE] - op_sccall));
    yc.constraint(lv[COL_TAPE_OPCODE] * (lv[COL_TAPE_OPCODE] - op_tload) * (one - lv[COL_FILTER_LOOKED]));
}
}  // namespace tape

namespace sccall {
enum : int {
    COL_SCCALL_TX_IDX = 0, COL_SCCALL_CALLER_ENV_IDX = 1, COL_SCCALL_CALLER_EXE_CTX = 2, COL_SCCALL_CALLER_CODE_CTX = 6, COL_SCCALL_CALLER_OP1_IMM = 10,
    COL_SCCALL_CLK_CALLER_CALL = 11, COL_SCCALL_CLK_CALLER_RET = 12, COL_SCCALL_CALLER_REG = 13, COL_SCCALL_CALLEE_ENV_IDX = 23,
    COL_SCCALL_CLK_CALLEE_END = 24, COL_SCCALL_IS_PADDING = 25, NUM_COL_SCCALL = 26
};
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    (void)nv;
    yc.constraint(lv[COL_SCCALL_CLK_CALLER_RET] - lv[COL_SCCALL_CLK_CALLER_CALL] - lv[COL_SCCALL_CALLER_OP1_IMM]);
}
}  // namespace sccall

// circuits/src/stark/lookup.rs:13-35 over generic rows
template <class T, class R, class C>
AIR_FN void eval_lookups_t(const R& lv, const R& nv, C& yc, int col_permuted_input, int col_permuted_table) {
    const T local_perm_input = lv[col_permuted_input];
    const T next_perm_table = nv[col_permuted_table];
    const T next_perm_input = nv[col_permuted_input];
    const T diff_input_prev = next_perm_input - local_perm_input;
    const T diff_input_table = next_perm_input - next_perm_table;
    yc.constraint(diff_input_prev * diff_input_table);
    yc.constraint_last_row(diff_input_table);
}

namespace program {
enum : int {
    COL_PROG_CODE_ADDR = 0, COL_PROG_PC = 4, COL_PROG_INST = 5, COL_PROG_COMP_PROG = 6, COL_PROG_COMP_PROG_PERM = 7, COL_PROG_EXEC_CODE_ADDR = 8,
    COL_PROG_EXEC_PC = 12, COL_PROG_EXEC_INST = 13, COL_PROG_EXEC_COMP_PROG = 14, COL_PROG_EXEC_COMP_PROG_PERM = 15, COL_PROG_FILTER_EXEC = 16,
    COL_PROG_FILTER_PROG_CHUNK = 17, NUM_PROG_COLS = 18
};
// beta = the table's compress challenge (ProgramStark::get_compress_challenge, program_stark.rs:49-58)
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc, const T beta) {
    const T b2 = beta * beta, b3 = b2 * beta;
    yc.constraint(lv[COL_PROG_CODE_ADDR] + lv[COL_PROG_CODE_ADDR + 1] * beta + lv[COL_PROG_CODE_ADDR + 2] * b2 + lv[COL_PROG_CODE_ADDR + 3] * b3 +
                  lv[COL_PROG_PC] * b2 * b2 + lv[COL_PROG_INST] * b2 * b3 - lv[COL_PROG_COMP_PROG]);
    yc.constraint(lv[COL_PROG_EXEC_CODE_ADDR] + lv[COL_PROG_EXEC_CODE_ADDR + 1] * beta + lv[COL_PROG_EXEC_CODE_ADDR + 2] * b2 +
                  lv[COL_PROG_EXEC_CODE_ADDR + 3] * b3 + lv[COL_PROG_EXEC_PC] * b2 * b2 + lv[COL_PROG_EXEC_INST] * b2 * b3 - lv[COL_PROG_EXEC_COMP_PROG]);
    eval_lookups_t<T, R, C>(lv, nv, yc, COL_PROG_EXEC_COMP_PROG_PERM, COL_PROG_COMP_PROG_PERM);
}
}  // namespace program

namespace prog_chunk {
enum : int {
    COL_PROG_CHUNK_CODE_ADDR = 0, COL_PROG_CHUNK_START_PC = 4, COL_PROG_CHUNK_INST = 5, COL_PROG_CHUNK_CAP = 13, COL_PROG_CHUNK_HASH = 17,
    COL_PROG_CHUNK_IS_FIRST_LINE = 29, COL_PROG_CHUNK_IS_RESULT_LINE = 30, COL_PROG_CHUNK_FILTER_LOOKING_PROG = 31, COL_PROG_CHUNK_IS_PADDING_LINE = 39,
    NUM_PROG_CHUNK_COLS = 40
};
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    const T one = kc<T>(1);
    const T lv_is_padding = lv[COL_PROG_CHUNK_IS_PADDING_LINE], nv_is_padding = nv[COL_PROG_CHUNK_IS_PADDING_LINE];
    const T lv_is_first_line = lv[COL_PROG_CHUNK_IS_FIRST_LINE], nv_is_first_line = nv[COL_PROG_CHUNK_IS_FIRST_LINE];
    const T lv_is_result_line = lv[COL_PROG_CHUNK_IS_RESULT_LINE];
    yc.constraint(lv_is_padding * (one - lv_is_padding));
    yc.constraint_transition((nv_is_padding - lv_is_padding) * (nv_is_padding - lv_is_padding - one));
    yc.constraint_first_row((one - lv_is_padding) * (one - lv_is_first_line));
    yc.constraint_transition((one - nv_is_padding) * (one - lv_is_result_line) * nv_is_first_line);
    yc.constraint_transition((one - nv_is_padding) * lv_is_result_line * (one - nv_is_first_line));
    for (int i = 0; i < 4; ++i)
        yc.constraint_transition((one - nv_is_padding) * (one - lv_is_result_line) * (nv[COL_PROG_CHUNK_CODE_ADDR + i] - lv[COL_PROG_CHUNK_CODE_ADDR + i]));
    yc.constraint(lv_is_first_line * lv[COL_PROG_CHUNK_START_PC]);
    yc.constraint_transition((one - nv_is_padding) * (one - lv_is_result_line) * (nv[COL_PROG_CHUNK_START_PC] - lv[COL_PROG_CHUNK_START_PC] - kc<T>(8)));
    for (int i = 0; i < 4; ++i) yc.constraint(lv_is_first_line * lv[COL_PROG_CHUNK_CAP + i]);
    for (int i = 0; i < 4; ++i) yc.constraint((one - nv_is_padding) * (one - nv_is_first_line) * (nv[COL_PROG_CHUNK_CAP + i] - lv[COL_PROG_CHUNK_HASH + 8 + i]));
    for (int i = 0; i < 8; ++i) {
        const T filter = lv[COL_PROG_CHUNK_FILTER_LOOKING_PROG + i];
        yc.constraint(filter * (one - filter));
        yc.constraint((one - lv_is_padding) * (one - lv_is_result_line) * (one - filter));
    }
    yc.constraint(lv_is_result_line * (one - lv[COL_PROG_CHUNK_FILTER_LOOKING_PROG]));
    for (int i = 0; i < 7; ++i) {
        const T after = lv[COL_PROG_CHUNK_FILTER_LOOKING_PROG + i], pre = lv[COL_PROG_CHUNK_FILTER_LOOKING_PROG + i + 1];
        yc.constraint(lv_is_result_line * (after - pre) * (one - (after - pre)));
    }
}
}  // namespace prog_chunk

namespace storage {
enum : int {
    COL_ST_ACCESS_IDX = 0, COL_ST_PRE_ROOT = 1, COL_ST_ROOT = 5, COL_ST_IS_WRITE = 9, COL_ST_LAYER = 10, COL_ST_LAYER_BIT = 11, COL_ST_ADDR_ACC = 12,
    COL_ST_ADDR = 13, COL_ST_PRE_PATH = 17, COL_ST_PATH = 21, COL_ST_SIB = 25, COL_ST_HASH_TYPE = 29, COL_ST_PRE_HASH = 30, COL_ST_HASH = 34,
    COL_ST_IS_LAYER_1 = 38, COL_ST_IS_LAYER_64 = 39, COL_ST_IS_LAYER_128 = 40, COL_ST_IS_LAYER_192 = 41, COL_ST_IS_LAYER_256 = 42,
    COL_ST_ACC_LAYER_MARKER = 43, COL_ST_FILTER_IS_HASH_BIT_0 = 44, COL_ST_FILTER_IS_HASH_BIT_1 = 45, COL_ST_FILTER_IS_FOR_PROG = 46,
    COL_ST_IS_PADDING = 47, NUM_COL_ST = 48
};
template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    const T one = kc<T>(1);
    const T lv_is_padding = lv[COL_ST_IS_PADDING], nv_is_padding = nv[COL_ST_IS_PADDING];
    const T lv_idx = lv[COL_ST_ACCESS_IDX], nv_idx = nv[COL_ST_ACCESS_IDX];
    const T lv_layer = lv[COL_ST_LAYER], nv_layer = nv[COL_ST_LAYER];
    const T didx = nv_idx - lv_idx;
    yc.constraint((one - lv_is_padding) * lv_is_padding);
    yc.constraint_transition((nv_is_padding - lv_is_padding) * (nv_is_padding - lv_is_padding - one));
    yc.constraint_first_row((one - lv_is_padding) * (lv_idx - one));
    yc.constraint_transition((one - nv_is_padding) * didx * (didx - one));
    yc.constraint_first_row((one - lv_is_padding) * (one - lv_layer));
    yc.constraint_transition((one - nv_is_padding) * (one - didx) * (nv_layer - lv_layer - one));
    yc.constraint_transition((one - nv_is_padding) * didx * (lv_layer - kc<T>(256)));
    yc.constraint_transition((one - nv_is_padding) * didx * (nv_layer - one));
    yc.constraint((one - nv_is_padding) * (lv_layer - kc<T>(256)) * (nv_layer - lv_layer - one));
    const int is_layer[5] = {COL_ST_IS_LAYER_1, COL_ST_IS_LAYER_64, COL_ST_IS_LAYER_128, COL_ST_IS_LAYER_192, COL_ST_IS_LAYER_256};
    const uint64_t layer_no[5] = {1, 64, 128, 192, 256};
    for (int i = 0; i < 5; ++i) yc.constraint(lv[is_layer[i]] * (one - lv[is_layer[i]]));
    yc.constraint_first_row((one - lv_is_padding) * (one - lv[COL_ST_IS_LAYER_1]));
    yc.constraint_transition((one - nv_is_padding) * didx * (one - nv[COL_ST_IS_LAYER_1]));
    for (int i = 0; i < 5; ++i) yc.constraint((lv[COL_ST_LAYER] - kc<T>(layer_no[i])) * lv[is_layer[i]]);
    yc.constraint_transition((one - nv_is_padding) * (one - didx) *
                             (nv[COL_ST_ACC_LAYER_MARKER] - lv[COL_ST_ACC_LAYER_MARKER] -
                              (nv[COL_ST_IS_LAYER_1] + nv[COL_ST_IS_LAYER_64] + nv[COL_ST_IS_LAYER_128] + nv[COL_ST_IS_LAYER_192] + nv[COL_ST_IS_LAYER_256])));
    yc.constraint_transition((one - nv_is_padding) * didx * (lv[COL_ST_ACC_LAYER_MARKER] - kc<T>(5)));
    yc.constraint_transition((one - nv_is_padding) * didx * (lv[COL_ST_HASH_TYPE] - one));
    yc.constraint_transition((one - nv_is_padding) * (one - didx) * lv[COL_ST_HASH_TYPE]);
    for (int i = 0; i < 4; ++i) yc.constraint(nv_is_padding * (nv[COL_ST_ROOT + i] - lv[COL_ST_ROOT + i]));
    for (int i = 0; i < 4; ++i) {
        yc.constraint_transition((one - nv_is_padding) * didx * (nv[COL_ST_PRE_ROOT + i] - lv[COL_ST_ROOT + i]));
        yc.constraint_transition((one - nv_is_padding) * (one - didx) * (nv[COL_ST_PRE_ROOT + i] - lv[COL_ST_PRE_ROOT + i]));
        yc.constraint_transition((one - nv_is_padding) * (one - didx) * (nv[COL_ST_ROOT + i] - lv[COL_ST_ROOT + i]));
        yc.constraint(lv[COL_ST_IS_LAYER_1] * (lv[COL_ST_PRE_ROOT + i] - lv[COL_ST_PRE_HASH + i]));
        yc.constraint(lv[COL_ST_IS_LAYER_1] * (lv[COL_ST_ROOT + i] - lv[COL_ST_HASH + i]));
    }
    yc.constraint(lv[COL_ST_LAYER_BIT] * (one - lv[COL_ST_LAYER_BIT]));
    yc.constraint_transition((one - lv[COL_ST_IS_LAYER_64] - lv[COL_ST_IS_LAYER_128] - lv[COL_ST_IS_LAYER_192] - lv[COL_ST_IS_LAYER_256]) *
                             (nv[COL_ST_ADDR_ACC] - lv[COL_ST_ADDR_ACC] * kc<T>(2) - nv[COL_ST_LAYER_BIT]));
    yc.constraint(lv[COL_ST_IS_LAYER_64] * (lv[COL_ST_ADDR_ACC] - lv[COL_ST_ADDR]));
    yc.constraint(lv[COL_ST_IS_LAYER_128] * (lv[COL_ST_ADDR_ACC] - lv[COL_ST_ADDR + 1]));
    yc.constraint(lv[COL_ST_IS_LAYER_192] * (lv[COL_ST_ADDR_ACC] - lv[COL_ST_ADDR + 2]));
    yc.constraint(lv[COL_ST_IS_LAYER_256] * (lv[COL_ST_ADDR_ACC] - lv[COL_ST_ADDR + 3]));
    for (int i = 0; i < 4; ++i) yc.constraint_transition((one - nv_is_padding) * (one - didx) * (lv[COL_ST_PATH + i] - nv[COL_ST_HASH + i]));
    yc.constraint((one - lv_is_padding) * (lv[COL_ST_FILTER_IS_HASH_BIT_0] + lv[COL_ST_LAYER_BIT] - one));
    yc.constraint((one - lv_is_padding) * (lv[COL_ST_FILTER_IS_HASH_BIT_1] - lv[COL_ST_LAYER_BIT]));
    yc.constraint(lv_is_padding * lv[COL_ST_FILTER_IS_HASH_BIT_0]);
    yc.constraint(lv_is_padding * lv[COL_ST_FILTER_IS_HASH_BIT_1]);
    yc.constraint(lv[COL_ST_FILTER_IS_FOR_PROG] * lv[COL_ST_IS_WRITE]);
    yc.constraint(lv[COL_ST_FILTER_IS_FOR_PROG] * (one - lv[COL_ST_IS_LAYER_256]));
}
}  // namespace storage

}  // namespace air
}  // namespace ola
