/* ORACLE (test infrastructure, NOT product code) -- AIRs of the StorageAccess, Program and ProgChunk tables, restated from
 * the Rust on their own (nothing under olavm_b200/ is included).  Constraints in source order.
 *   StorageAccess  circuits/src/builtins/storage/{columns.rs:4-44, storage_access_stark.rs:145-373}
 *   Program        circuits/src/program/{columns.rs:3-16, program_stark.rs:60-105}   (compress challenge beta)
 *   ProgChunk      circuits/src/program/{columns.rs:47-62, prog_chunk_stark.rs:121-216} */
#ifndef ORC_AIR_STORAGE_PROGRAM_HPP
#define ORC_AIR_STORAGE_PROGRAM_HPP
#include "air_builtins.hpp"

namespace orc {

namespace storage_air {
enum {
    ACCESS_IDX = 0,
    PRE_ROOT = ACCESS_IDX + 1, /* 4 */
    ROOT = PRE_ROOT + 4,       /* 4 */
    IS_WRITE = ROOT + 4, LAYER, LAYER_BIT, ADDR_ACC,
    ADDR = ADDR_ACC + 1,  /* 4 */
    PRE_PATH = ADDR + 4,  /* 4 */
    PATH = PRE_PATH + 4,  /* 4 */
    SIB = PATH + 4,       /* 4 */
    HASH_TYPE = SIB + 4,
    PRE_HASH = HASH_TYPE + 1, /* 4 */
    HASH = PRE_HASH + 4,      /* 4 */
    IS_LAYER_1 = HASH + 4, IS_LAYER_64, IS_LAYER_128, IS_LAYER_192, IS_LAYER_256, ACC_LAYER_MARKER, FILTER_IS_HASH_BIT_0, FILTER_IS_HASH_BIT_1,
    FILTER_IS_FOR_PROG, IS_PADDING, NUM_COLS
};
static_assert(NUM_COLS == 48, "storage/columns.rs layout");

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    const T ONE = T::one();
    const T lv_is_padding = lv[IS_PADDING], nv_is_padding = nv[IS_PADDING];
    const T lv_idx = lv[ACCESS_IDX], nv_idx = nv[ACCESS_IDX];
    const T lv_layer = lv[LAYER], nv_layer = nv[LAYER];
    const T d_idx = nv_idx - lv_idx;

    yc.constraint((ONE - lv_is_padding) * lv_is_padding);
    yc.constraint_transition((nv_is_padding - lv_is_padding) * (nv_is_padding - lv_is_padding - ONE));
    yc.constraint_first_row((ONE - lv_is_padding) * (lv_idx - ONE));
    yc.constraint_transition((ONE - nv_is_padding) * d_idx * (d_idx - ONE));

    yc.constraint_first_row((ONE - lv_is_padding) * (ONE - lv_layer));
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - d_idx) * (nv_layer - lv_layer - ONE));
    yc.constraint_transition((ONE - nv_is_padding) * d_idx * (lv_layer - T::c(256)));
    yc.constraint_transition((ONE - nv_is_padding) * d_idx * (nv_layer - ONE));
    yc.constraint((ONE - nv_is_padding) * (lv_layer - T::c(256)) * (nv_layer - lv_layer - ONE));

    const int marks[5] = {IS_LAYER_1, IS_LAYER_64, IS_LAYER_128, IS_LAYER_192, IS_LAYER_256};
    const uint64_t at[5] = {1, 64, 128, 192, 256};
    for (int i = 0; i < 5; i++) yc.constraint(lv[marks[i]] * (ONE - lv[marks[i]]));
    yc.constraint_first_row((ONE - lv_is_padding) * (ONE - lv[IS_LAYER_1]));
    yc.constraint_transition((ONE - nv_is_padding) * d_idx * (ONE - nv[IS_LAYER_1]));
    yc.constraint((lv[LAYER] - ONE) * lv[IS_LAYER_1]);
    for (int i = 1; i < 5; i++) yc.constraint((lv[LAYER] - T::c(at[i])) * lv[marks[i]]);
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - d_idx) *
                             (nv[ACC_LAYER_MARKER] - lv[ACC_LAYER_MARKER] -
                              (nv[IS_LAYER_1] + nv[IS_LAYER_64] + nv[IS_LAYER_128] + nv[IS_LAYER_192] + nv[IS_LAYER_256])));
    yc.constraint_transition((ONE - nv_is_padding) * d_idx * (lv[ACC_LAYER_MARKER] - T::c(5)));

    yc.constraint_transition((ONE - nv_is_padding) * d_idx * (lv[HASH_TYPE] - ONE));
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - d_idx) * lv[HASH_TYPE]);

    for (int i = 0; i < 4; i++) yc.constraint(nv_is_padding * (nv[ROOT + i] - lv[ROOT + i]));
    for (int i = 0; i < 4; i++) {
        yc.constraint_transition((ONE - nv_is_padding) * d_idx * (nv[PRE_ROOT + i] - lv[ROOT + i]));
        yc.constraint_transition((ONE - nv_is_padding) * (ONE - d_idx) * (nv[PRE_ROOT + i] - lv[PRE_ROOT + i]));
        yc.constraint_transition((ONE - nv_is_padding) * (ONE - d_idx) * (nv[ROOT + i] - lv[ROOT + i]));
        yc.constraint(lv[IS_LAYER_1] * (lv[PRE_ROOT + i] - lv[PRE_HASH + i]));
        yc.constraint(lv[IS_LAYER_1] * (lv[ROOT + i] - lv[HASH + i]));
    }

    yc.constraint(lv[LAYER_BIT] * (ONE - lv[LAYER_BIT]));
    yc.constraint_transition((ONE - lv[IS_LAYER_64] - lv[IS_LAYER_128] - lv[IS_LAYER_192] - lv[IS_LAYER_256]) *
                             (nv[ADDR_ACC] - lv[ADDR_ACC] * T::c(2) - nv[LAYER_BIT]));
    for (int i = 0; i < 4; i++) yc.constraint(lv[marks[i + 1]] * (lv[ADDR_ACC] - lv[ADDR + i]));

    for (int i = 0; i < 4; i++) yc.constraint_transition((ONE - nv_is_padding) * (ONE - d_idx) * (lv[PATH + i] - nv[HASH + i]));

    yc.constraint((ONE - lv_is_padding) * (lv[FILTER_IS_HASH_BIT_0] + lv[LAYER_BIT] - ONE));
    yc.constraint((ONE - lv_is_padding) * (lv[FILTER_IS_HASH_BIT_1] - lv[LAYER_BIT]));
    yc.constraint(lv_is_padding * lv[FILTER_IS_HASH_BIT_0]);
    yc.constraint(lv_is_padding * lv[FILTER_IS_HASH_BIT_1]);
    yc.constraint(lv[FILTER_IS_FOR_PROG] * lv[IS_WRITE]);
    yc.constraint(lv[FILTER_IS_FOR_PROG] * (ONE - lv[IS_LAYER_256]));
}
}  // namespace storage_air

namespace program_air {
enum {
    CODE_ADDR = 0, /* 4 */
    PC = 4, INST, COMP_PROG, COMP_PROG_PERM,
    EXEC_CODE_ADDR = COMP_PROG_PERM + 1, /* 4 */
    EXEC_PC = EXEC_CODE_ADDR + 4, EXEC_INST, EXEC_COMP_PROG, EXEC_COMP_PROG_PERM, FILTER_EXEC, FILTER_PROG_CHUNK, NUM_COLS
};
static_assert(NUM_COLS == 18, "program/columns.rs layout");

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc, P<O> beta) {
    typedef P<O> T;
    const T b2 = beta * beta, b3 = b2 * beta; /* beta.square(), beta.cube() */
    yc.constraint(lv[CODE_ADDR] + lv[CODE_ADDR + 1] * beta + lv[CODE_ADDR + 2] * b2 + lv[CODE_ADDR + 3] * b3 + lv[PC] * b2 * b2 + lv[INST] * b2 * b3 -
                  lv[COMP_PROG]);
    yc.constraint(lv[EXEC_CODE_ADDR] + lv[EXEC_CODE_ADDR + 1] * beta + lv[EXEC_CODE_ADDR + 2] * b2 + lv[EXEC_CODE_ADDR + 3] * b3 + lv[EXEC_PC] * b2 * b2 +
                  lv[EXEC_INST] * b2 * b3 - lv[EXEC_COMP_PROG]);
    air_eval_lookups<O>(lv, nv, yc, EXEC_COMP_PROG_PERM, COMP_PROG_PERM);
}
}  // namespace program_air

namespace prog_chunk_air {
enum {
    CODE_ADDR = 0, /* 4 */
    START_PC = 4,
    INST = START_PC + 1, /* 8 */
    CAP = INST + 8,      /* 4 */
    HASH = CAP + 4,      /* 12 */
    IS_FIRST_LINE = HASH + 12, IS_RESULT_LINE,
    FILTER_LOOKING_PROG = IS_RESULT_LINE + 1, /* 8 */
    IS_PADDING_LINE = FILTER_LOOKING_PROG + 8, NUM_COLS
};
static_assert(NUM_COLS == 40, "program/columns.rs chunk layout");

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    const T ONE = T::one();
    const T lv_is_padding = lv[IS_PADDING_LINE], nv_is_padding = nv[IS_PADDING_LINE];
    const T lv_is_first_line = lv[IS_FIRST_LINE], nv_is_first_line = nv[IS_FIRST_LINE];
    const T lv_is_result_line = lv[IS_RESULT_LINE];
    yc.constraint(lv_is_padding * (ONE - lv_is_padding));
    yc.constraint_transition((nv_is_padding - lv_is_padding) * (nv_is_padding - lv_is_padding - ONE));
    yc.constraint_first_row((ONE - lv_is_padding) * (ONE - lv_is_first_line));
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - lv_is_result_line) * nv_is_first_line);
    yc.constraint_transition((ONE - nv_is_padding) * lv_is_result_line * (ONE - nv_is_first_line));
    for (int i = 0; i < 4; i++) yc.constraint_transition((ONE - nv_is_padding) * (ONE - lv_is_result_line) * (nv[CODE_ADDR + i] - lv[CODE_ADDR + i]));
    yc.constraint(lv_is_first_line * lv[START_PC]);
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - lv_is_result_line) * (nv[START_PC] - lv[START_PC] - T::c(8)));
    for (int i = 0; i < 4; i++) yc.constraint(lv_is_first_line * lv[CAP + i]);
    for (int i = 0; i < 4; i++) yc.constraint((ONE - nv_is_padding) * (ONE - nv_is_first_line) * (nv[CAP + i] - lv[HASH + 8 + i]));
    for (int i = 0; i < 8; i++) {
        const T filter = lv[FILTER_LOOKING_PROG + i];
        yc.constraint(filter * (ONE - filter));
        yc.constraint((ONE - lv_is_padding) * (ONE - lv_is_result_line) * (ONE - filter));
    }
    yc.constraint(lv_is_result_line * (ONE - lv[FILTER_LOOKING_PROG]));
    for (int i = 0; i < 7; i++) {
        const T after = lv[FILTER_LOOKING_PROG + i], pre = lv[FILTER_LOOKING_PROG + i + 1];
        yc.constraint(lv_is_result_line * (after - pre) * (ONE - (after - pre)));
    }
}
}  // namespace prog_chunk_air

}  // namespace orc
#endif
