/* ORACLE (test infrastructure, NOT product code) -- the CPU table's AIR, restated from the Rust on its own (nothing under
 * olavm_b200/ is included): circuits/src/cpu/cpu_stark.rs:330-946 (CpuStark::eval_packed_generic and its helper
 * functions, CpuAdjacentRowWrapper::from_vars) and the per-opcode files circuits/src/cpu/{simple_arithmatic_op, mov, call,
 * ret, mload, mstore, storage, tape, call_sc}.rs.  Column layout: circuits/src/cpu/columns.rs:4-133 (derived below exactly as
 * the Rust derives it: each index is the previous one plus a width).  Opcode masks: core/src/vm/opcodes.rs:76-109
 * (1 << bit).  Constraints are emitted in source order. */
#ifndef ORC_AIR_CPU_HPP
#define ORC_AIR_CPU_HPP
#include "stark.hpp"

namespace orc {
namespace cpu_air {

enum { REGISTER_NUM = 10, CTX_REGISTER_NUM = 4 }; /* core/src/program/mod.rs */
/* columns.rs:4-133 */
enum {
    TX_IDX = 0,
    ENV_IDX = TX_IDX + 1,
    CALL_SC_CNT = ENV_IDX + 1,
    ADDR_STORAGE = CALL_SC_CNT + 1,              /* range of CTX_REGISTER_NUM */
    ADDR_CODE = ADDR_STORAGE + CTX_REGISTER_NUM, /* range of CTX_REGISTER_NUM */
    TP = ADDR_CODE + CTX_REGISTER_NUM,
    CLK = TP + 1,
    PC = CLK + 1,
    IS_EXT_LINE = PC + 1,
    EXT_CNT = IS_EXT_LINE + 1,
    REGS = EXT_CNT + 1, /* range of REGISTER_NUM */
    INST = REGS + REGISTER_NUM,
    OP1_IMM = INST + 1,
    OPCODE = OP1_IMM + 1,
    IMM_VAL = OPCODE + 1,
    OP0 = IMM_VAL + 1,
    OP1 = OP0 + 1,
    DST = OP1 + 1,
    AUX0 = DST + 1,
    AUX1 = AUX0 + 1,
    IDX_STORAGE = AUX1 + 1,
    S_OP0 = IDX_STORAGE + 1, /* three register-selector ranges */
    S_OP1 = S_OP0 + REGISTER_NUM,
    S_DST = S_OP1 + REGISTER_NUM,
    S_SIMPLE_ARITHMATIC_OP = S_DST + REGISTER_NUM,
    S_MOV, S_JMP, S_CJMP, S_CALL, S_RET, S_MLOAD, S_MSTORE, S_END, S_RC, S_BITWISE, S_NOT, S_GTE, S_PSDN, S_SLOAD, S_SSTORE, S_TLOAD,
    S_TSTORE, S_CALL_SC,
    NUM_OP_SELECTOR = S_CALL_SC - S_SIMPLE_ARITHMATIC_OP + 1,
    IS_ENTRY_SC = S_CALL_SC + 1,
    IS_NEXT_LINE_DIFF_INST,
    IS_NEXT_LINE_SAME_TX,
    FILTER_TAPE_LOOKING,
    IS_SCCALL_EXT_LINE,
    IS_STORAGE_EXT_LINE,
    FILTER_SCCALL_END,
    FILTER_LOOKING_PROG_IMM,
    IS_PADDING,
    NUM_COLS
};
static_assert(NUM_COLS == 94 && NUM_OP_SELECTOR == 19, "columns.rs layout");

/* OlaOpcode::binary_bit_mask: 1 << binary_bit_shift */
inline uint64_t mask(int shift) { return (uint64_t)1 << shift; }
enum { B_ADD = 31, B_MUL = 30, B_EQ = 29, B_ASSERT = 28, B_MOV = 27, B_JMP = 26, B_CJMP = 25, B_CALL = 24, B_RET = 23, B_MLOAD = 22,
       B_MSTORE = 21, B_END = 20, B_RC = 19, B_AND = 18, B_OR = 17, B_XOR = 16, B_NOT = 15, B_NEQ = 14, B_GTE = 13, B_POSEIDON = 12,
       B_SLOAD = 11, B_SSTORE = 10, B_TLOAD = 9, B_TSTORE = 8, B_SCCALL = 7 };

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    const T ONE = T::one();
    /* ---- CpuAdjacentRowWrapper::from_vars ---- */
    const T lv_is_padding = lv[IS_PADDING], nv_is_padding = nv[IS_PADDING];
    const T lv_is_ext_inst = lv[S_SLOAD] + lv[S_SSTORE] + lv[S_TLOAD] + lv[S_TSTORE] + lv[S_CALL_SC] + lv[S_END];
    const T nv_is_ext_inst = nv[S_SLOAD] + nv[S_SSTORE] + nv[S_TLOAD] + nv[S_TSTORE] + nv[S_CALL_SC] + nv[S_END];
    const T lv_is_entry_sc = lv[IS_ENTRY_SC];
    const T lv_ext_length = lv[S_SLOAD] + lv[S_SSTORE] + lv[S_TLOAD] * (lv[OP0] * lv[OP1] + (ONE - lv[OP0])) + lv[S_TSTORE] * lv[OP1] +
                            lv[S_CALL_SC] + lv[S_END] * (ONE - lv_is_entry_sc);
    const T is_crossing_inst = lv[IS_NEXT_LINE_DIFF_INST];
    const T is_in_same_tx = lv[IS_NEXT_LINE_SAME_TX];

    /* ---- constraint_wrapper_cols ---- */
    yc.constraint(lv_is_padding * (lv_is_padding - ONE));
    yc.constraint_transition((nv_is_padding - lv_is_padding) * (nv_is_padding - lv_is_padding - ONE));
    yc.constraint(lv_is_padding * (lv[S_END] - ONE));
    yc.constraint(lv_is_entry_sc * nv[ENV_IDX]);
    yc.constraint((ONE - nv_is_padding) * is_in_same_tx * (nv[TX_IDX] - lv[TX_IDX]));
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - is_in_same_tx) * (nv[TX_IDX] - lv[TX_IDX] - ONE));
    yc.constraint(is_crossing_inst * (lv_ext_length - lv[EXT_CNT]));

    /* ---- constraint_tx_init ---- */
    yc.constraint_first_row(lv[TX_IDX]);
    yc.constraint_first_row(lv[ENV_IDX]);
    yc.constraint_first_row(lv[CALL_SC_CNT]);
    yc.constraint_first_row(lv[CLK]);
    yc.constraint_first_row(lv[PC]);
    for (int r = REGS; r < REGS + REGISTER_NUM; r++) yc.constraint_first_row(lv[r]);
    yc.constraint_transition(is_in_same_tx * (nv[TX_IDX] - lv[TX_IDX]));
    yc.constraint_transition((ONE - is_in_same_tx) * nv[ENV_IDX]);
    yc.constraint_transition((ONE - is_in_same_tx) * nv[CALL_SC_CNT]);
    yc.constraint_transition((ONE - is_in_same_tx) * nv[TP]);
    yc.constraint_transition((ONE - is_in_same_tx) * nv[CLK]);
    yc.constraint_transition((ONE - is_in_same_tx) * nv[PC]);
    for (int r = REGS; r < REGS + REGISTER_NUM; r++) yc.constraint_transition((ONE - is_in_same_tx) * nv[r]);

    /* ---- inline part of eval_packed_generic ---- */
    yc.constraint_transition((ONE - nv_is_padding) * (ONE - lv[S_END]) * (nv[TX_IDX] - lv[TX_IDX]));
    yc.constraint_transition((ONE - nv_is_padding) * lv_is_entry_sc * lv[S_END] * (nv[TX_IDX] - lv[TX_IDX] - ONE));
    for (int i = 0; i < CTX_REGISTER_NUM; i++) {
        yc.constraint_transition((ONE - nv_is_padding) * (ONE - lv[S_END]) * (ONE - lv[S_CALL_SC]) * (nv[ADDR_STORAGE + i] - lv[ADDR_STORAGE + i]));
        yc.constraint_transition((ONE - nv_is_padding) * (ONE - lv[S_END]) * (ONE - lv[S_CALL_SC]) * (nv[ADDR_CODE + i] - lv[ADDR_CODE + i]));
    }
    yc.constraint((ONE - lv[IS_PADDING] - lv[IS_EXT_LINE]) * lv[OP1_IMM] * (ONE - lv[FILTER_LOOKING_PROG_IMM]));
    yc.constraint((ONE - lv[IS_PADDING] - lv[IS_EXT_LINE]) * (lv[S_MLOAD] + lv[S_MSTORE]) * (ONE - lv[FILTER_LOOKING_PROG_IMM]));

    /* ---- constraint_ext_lines ---- */
    yc.constraint((ONE - lv_is_ext_inst) * lv[IS_EXT_LINE]);
    yc.constraint(lv_is_ext_inst * (lv_ext_length - lv[EXT_CNT]) * (ONE - nv[IS_EXT_LINE]));
    yc.constraint(lv_is_ext_inst * (ONE - lv[IS_EXT_LINE]) * lv[EXT_CNT]);
    yc.constraint(nv_is_ext_inst * nv[IS_EXT_LINE] * (nv[EXT_CNT] - lv[EXT_CNT] - ONE));
    yc.constraint(nv[IS_EXT_LINE] * (nv[OPCODE] - lv[OPCODE]));
    for (int s = S_SIMPLE_ARITHMATIC_OP; s < S_SIMPLE_ARITHMATIC_OP + NUM_OP_SELECTOR; s++) yc.constraint(nv[IS_EXT_LINE] * (nv[s] - lv[s]));
    yc.constraint(nv[IS_EXT_LINE] * (nv[OP1_IMM] - lv[OP1_IMM]));

    /* ---- constraint_env_idx ---- */
    yc.constraint_transition(lv[S_CALL_SC] * is_crossing_inst * (nv[CALL_SC_CNT] - lv[CALL_SC_CNT] - ONE));
    yc.constraint_transition(is_in_same_tx * (ONE - lv[S_CALL_SC]) * (nv[CALL_SC_CNT] - lv[CALL_SC_CNT]));
    yc.constraint(lv[S_CALL_SC] * (ONE - is_crossing_inst) * (nv[CALL_SC_CNT] - lv[CALL_SC_CNT]));
    yc.constraint(lv[S_CALL_SC] * is_crossing_inst * (nv[ENV_IDX] - lv[CALL_SC_CNT]));
    yc.constraint((ONE - lv[S_CALL_SC] - lv[S_END]) * (nv[ENV_IDX] - lv[ENV_IDX]));
    yc.constraint(lv[S_CALL_SC] * (ONE - is_crossing_inst) * (nv[ENV_IDX] - lv[ENV_IDX]));
    yc.constraint(lv[S_END] * lv[IS_EXT_LINE] * (nv[ENV_IDX] - lv[ENV_IDX]));

    /* ---- constraint_opcode_selector ---- */
    {
        struct SelOp { int col; uint64_t op; };
        const SelOp ops_to_op[NUM_OP_SELECTOR] = {
            {S_SIMPLE_ARITHMATIC_OP, 0}, {S_MOV, mask(B_MOV)}, {S_JMP, mask(B_JMP)}, {S_CJMP, mask(B_CJMP)}, {S_CALL, mask(B_CALL)},
            {S_RET, mask(B_RET)}, {S_MLOAD, mask(B_MLOAD)}, {S_MSTORE, mask(B_MSTORE)}, {S_END, mask(B_END)}, {S_RC, mask(B_RC)},
            {S_BITWISE, 0}, {S_NOT, mask(B_NOT)}, {S_GTE, mask(B_GTE)}, {S_PSDN, mask(B_POSEIDON)}, {S_SLOAD, mask(B_SLOAD)},
            {S_SSTORE, mask(B_SSTORE)}, {S_TLOAD, mask(B_TLOAD)}, {S_TSTORE, mask(B_TSTORE)}, {S_CALL_SC, mask(B_SCCALL)}};
        yc.constraint(lv[S_SIMPLE_ARITHMATIC_OP] * (lv[OPCODE] - T::c(mask(B_ADD))) * (lv[OPCODE] - T::c(mask(B_MUL))) * (lv[OPCODE] - T::c(mask(B_EQ))) *
                      (lv[OPCODE] - T::c(mask(B_NEQ))) * (lv[OPCODE] - T::c(mask(B_ASSERT))));
        yc.constraint(lv[S_BITWISE] * (lv[OPCODE] - T::c(mask(B_AND))) * (lv[OPCODE] - T::c(mask(B_OR))) * (lv[OPCODE] - T::c(mask(B_XOR))));
        for (const SelOp& so : ops_to_op) yc.constraint(lv[so.col] * (ONE - lv[so.col]));
        T sum_s_op = T::zero();
        for (const SelOp& so : ops_to_op) sum_s_op = sum_s_op + lv[so.col];
        yc.constraint(ONE - sum_s_op);
        T cal_opcode = T::zero();
        for (const SelOp& so : ops_to_op) cal_opcode = cal_opcode + lv[so.col] * T::c(so.op);
        yc.constraint((lv[OPCODE] - cal_opcode) * (ONE - lv[S_BITWISE] - lv[S_SIMPLE_ARITHMATIC_OP]));
    }

    /* ---- constraint_instruction_encode: OP1_IMM_SHIFT 62, OP0 / OP1 / DST shifts start at 61 / 51 / 41 and halve while
     * walking the selectors in REVERSE (r9 first) ---- */
    {
        yc.constraint(lv[OP1_IMM] * (ONE - lv[OP1_IMM]));
        T instruction = lv[OP1_IMM] * T::c((uint64_t)1 << 62);
        const int starts[3] = {61, 51, 41}, bases[3] = {S_OP0, S_OP1, S_DST};
        for (int g = 0; g < 3; g++)
            for (int index = 0; index < REGISTER_NUM; index++) {
                const uint64_t shift = ((uint64_t)1 << starts[g]) / ((uint64_t)1 << index);
                instruction = instruction + lv[bases[g] + (REGISTER_NUM - 1 - index)] * T::c(shift);
            }
        instruction = instruction + lv[OPCODE];
        yc.constraint((ONE - lv[IS_EXT_LINE]) * (lv[INST] - instruction));
        yc.constraint((ONE - lv[IS_EXT_LINE]) * (lv[OP1_IMM] * (lv[OP1] - lv[IMM_VAL])));
    }

    /* ---- constraint_operands_mathches_registers ---- */
    {
        const int bases[3] = {S_OP0, S_OP1, S_DST};
        for (int g = 0; g < 3; g++)
            for (int i = 0; i < REGISTER_NUM; i++) yc.constraint((ONE - lv[IS_EXT_LINE]) * lv[bases[g] + i] * (ONE - lv[bases[g] + i]));
        T sums[3];
        for (int g = 0; g < 3; g++) {
            sums[g] = T::zero();
            for (int i = 0; i < REGISTER_NUM; i++) sums[g] = sums[g] + lv[bases[g] + i];
            yc.constraint((ONE - lv[IS_EXT_LINE]) * sums[g] * (ONE - sums[g]));
        }
        T op0_sum = T::zero(), op1_sum = T::zero(), dst_sum = T::zero();
        for (int i = 0; i < REGISTER_NUM; i++) op0_sum = op0_sum + lv[S_OP0 + i] * lv[REGS + i];
        yc.constraint((ONE - lv[IS_EXT_LINE]) * sums[0] * (lv[OP0] - op0_sum));
        for (int i = 0; i < REGISTER_NUM; i++) op1_sum = op1_sum + lv[S_OP1 + i] * lv[REGS + i];
        yc.constraint((ONE - lv[IS_EXT_LINE]) * sums[1] * (lv[OP1] - op1_sum));
        for (int i = 0; i < REGISTER_NUM; i++) dst_sum = dst_sum + lv[S_DST + i] * nv[REGS + i];
        yc.constraint((ONE - lv[IS_EXT_LINE]) * sums[2] * (lv[DST] - dst_sum));
    }

    /* ---- constraint_env_unchanged_clk ---- */
    yc.constraint(nv[IS_EXT_LINE] * (ONE - nv[S_END]) * (nv[CLK] - lv[CLK]));
    yc.constraint(is_in_same_tx * (ONE - lv[S_CALL_SC] - lv[S_END]) * (ONE - nv[IS_EXT_LINE]) * (nv[CLK] - lv[CLK] - ONE));

    /* ---- constraint_env_unchanged_pc (its first constraint is the clk one again) ---- */
    {
        yc.constraint(nv[IS_EXT_LINE] * (ONE - nv[S_END]) * (nv[CLK] - lv[CLK]));
        const T instruction_size = (ONE - lv[S_MLOAD] - lv[S_MSTORE]) * (ONE + lv[OP1_IMM]) + (lv[S_MLOAD] + lv[S_MSTORE]) * T::c(2);
        const T pc_incr = (ONE - (lv[S_JMP] + lv[S_CJMP] + lv[S_CALL] + lv[S_RET])) * (lv[PC] + instruction_size);
        const T pc_jmp = lv[S_JMP] * lv[OP1];
        const T pc_cjmp = lv[S_CJMP] * ((ONE - lv[OP0]) * (lv[PC] + instruction_size) + lv[OP0] * lv[OP1]);
        const T pc_call = lv[S_CALL] * lv[OP1];
        const T pc_ret = lv[S_RET] * lv[DST];
        yc.constraint((ONE - nv[IS_EXT_LINE]) * (ONE - lv[S_END] - lv[S_CALL_SC]) * (nv[PC] - (pc_incr + pc_jmp + pc_cjmp + pc_call + pc_ret)));
        yc.constraint((ONE - nv[IS_EXT_LINE]) * lv[S_CJMP] * lv[OP0] * (ONE - lv[OP0]));
    }

    /* ---- constraint_reg_consistency ---- */
    {
        const T multi_reg_change = lv[S_SLOAD] + lv[S_PSDN] + lv[S_CALL_SC] * is_crossing_inst + lv[S_END] * (ONE - lv[IS_EXT_LINE]);
        for (int i = 0; i < REGISTER_NUM - 1; i++)
            yc.constraint_transition((ONE - multi_reg_change) * (ONE - lv[S_DST + i]) * (nv[REGS + i] - lv[REGS + i]));
        const int fp = REGISTER_NUM - 1;
        yc.constraint_transition((ONE - lv[S_RET] - lv[S_CALL_SC] * is_crossing_inst - lv[S_END]) * (ONE - lv[S_DST + fp]) * (nv[REGS + fp] - lv[REGS + fp]));
    }

    /* ---- simple_arithmatic_op.rs ---- */
    {
        const T op = lv[OPCODE], s = lv[S_SIMPLE_ARITHMATIC_OP];
        const T m_add = T::c(mask(B_ADD)), m_mul = T::c(mask(B_MUL)), m_eq = T::c(mask(B_EQ)), m_neq = T::c(mask(B_NEQ)), m_assert = T::c(mask(B_ASSERT));
        const T is_add = s * (op - m_mul) * (op - m_eq) * (op - m_neq) * (op - m_assert);
        const T is_mul = s * (op - m_add) * (op - m_eq) * (op - m_neq) * (op - m_assert);
        const T is_eq = s * (op - m_add) * (op - m_mul) * (op - m_neq) * (op - m_assert);
        const T is_neq = s * (op - m_add) * (op - m_mul) * (op - m_eq) * (op - m_assert);
        const T is_assert = s * (op - m_add) * (op - m_mul) * (op - m_eq) * (op - m_neq);
        yc.constraint(is_add * (lv[DST] - (lv[OP0] + lv[OP1])));
        yc.constraint(is_mul * (lv[DST] - lv[OP0] * lv[OP1]));
        const T op_diff = lv[OP0] - lv[OP1];
        const T diff_aux = op_diff * lv[AUX0];
        const T res = lv[DST];
        const T eq_cs = is_eq * (res * op_diff + (ONE - res) * (ONE - diff_aux));
        const T neq_cs = is_neq * ((ONE - res) * op_diff + res * (ONE - diff_aux));
        yc.constraint(eq_cs + neq_cs);
        yc.constraint(is_assert * (ONE - lv[OP1]));
    }
    /* ---- mov.rs ---- */
    yc.constraint(lv[S_MOV] * (lv[DST] - lv[OP1]));
    /* ---- call.rs ---- */
    {
        const T two = ONE + ONE;
        const T fp = lv[REGS + REGISTER_NUM - 1];
        const T op0_cs = lv[OP0] + ONE - fp;
        const T op1_cs = lv[OP1_IMM] * (lv[DST] - lv[PC] - two) + (ONE - lv[OP1_IMM]) * (lv[DST] - lv[PC] - ONE);
        const T aux0_cs = lv[AUX0] - fp + two;
        yc.constraint(lv[S_CALL] * (op0_cs + op1_cs + aux0_cs));
    }
    /* ---- ret.rs ---- */
    {
        const T fp = lv[REGS + REGISTER_NUM - 1];
        const T op0_cs = lv[OP0] + ONE - fp;
        const T dst_cs = lv[DST] - nv[PC];
        const T aux0_cs = lv[AUX0] + ONE + ONE - fp;
        yc.constraint(lv[S_RET] * (op0_cs + dst_cs + aux0_cs));
        yc.constraint_transition(lv[S_RET] * (nv[REGS + REGISTER_NUM - 1] - lv[AUX1]));
    }
    /* ---- mload.rs, mstore.rs ---- */
    {
        const int sel[2] = {S_MLOAD, S_MSTORE};
        for (int k = 0; k < 2; k++) {
            yc.constraint(lv[sel[k]] * (ONE - lv[OP1_IMM]) * (lv[AUX0] - lv[IMM_VAL]));
            yc.constraint(lv[sel[k]] * lv[OP1_IMM] * (lv[AUX1] - lv[OP0] - lv[OP1]));
            yc.constraint(lv[sel[k]] * (ONE - lv[OP1_IMM]) * (lv[AUX1] - lv[OP0] - lv[AUX0] * lv[OP1]));
        }
    }
    /* ---- storage.rs ---- */
    {
        const T is_storage_op = lv[S_SSTORE] + lv[S_SLOAD];
        yc.constraint_first_row(lv[IDX_STORAGE] - is_storage_op);
        yc.constraint_transition(nv[IDX_STORAGE] - lv[IDX_STORAGE] - nv[IS_STORAGE_EXT_LINE]);
        yc.constraint(is_storage_op * (ONE - lv[IS_EXT_LINE]) * (nv[OP0] - lv[OP0]));
        yc.constraint(is_storage_op * (ONE - lv[IS_EXT_LINE]) * (nv[OP1] - lv[OP1]));
        yc.constraint(is_storage_op * lv[IS_EXT_LINE] * (lv[S_OP0] - lv[OP0]));
        for (int i = 1; i <= 3; i++) yc.constraint(is_storage_op * lv[IS_EXT_LINE] * (lv[S_OP0 + i] - lv[S_OP0 + i - 1] - ONE));
        yc.constraint(is_storage_op * lv[IS_EXT_LINE] * (lv[S_OP1] - lv[OP1]));
        for (int i = 1; i <= 3; i++) yc.constraint(is_storage_op * lv[IS_EXT_LINE] * (lv[S_OP1 + i] - lv[S_OP1 + i - 1] - ONE));
        yc.constraint(is_storage_op * lv[IS_EXT_LINE] * (ONE - lv[IS_STORAGE_EXT_LINE]));
        yc.constraint((ONE - is_storage_op) * lv[IS_STORAGE_EXT_LINE]);
        yc.constraint(is_storage_op * (ONE - lv[IS_EXT_LINE]) * lv[IS_STORAGE_EXT_LINE]);
    }
    /* ---- tape.rs ---- */
    {
        yc.constraint((nv[S_TSTORE] + nv[S_TLOAD]) * nv[IS_EXT_LINE] * (nv[OP0] - lv[OP0]));
        yc.constraint((nv[S_TSTORE] + nv[S_TLOAD]) * nv[IS_EXT_LINE] * (nv[OP1] - lv[OP1]));
        yc.constraint((lv[S_TSTORE] + lv[S_TLOAD]) * lv[IS_EXT_LINE] * nv[IS_EXT_LINE] * (nv[AUX0] - lv[AUX0] - ONE));
        yc.constraint(lv[S_TSTORE] * (ONE - lv[IS_EXT_LINE]) * (lv[TP] - nv[S_OP0]));
        yc.constraint(lv[S_TSTORE] * lv[IS_EXT_LINE] * nv[IS_EXT_LINE] * (nv[S_OP0] - lv[S_OP0] - ONE));
        yc.constraint(lv[S_TSTORE] * (ONE - nv[IS_EXT_LINE]) * (nv[TP] - lv[S_OP0] - ONE));
        yc.constraint(lv[S_TLOAD] * lv[OP0] * (ONE - lv[IS_EXT_LINE]) * (nv[S_OP0] + lv[OP1] - lv[TP]));
        yc.constraint(lv[S_TLOAD] * (ONE - lv[OP0]) * (ONE - lv[IS_EXT_LINE]) * (nv[S_OP0] - lv[OP1]));
        yc.constraint((lv[S_TSTORE] + lv[S_TLOAD]) * lv[IS_EXT_LINE] * nv[IS_EXT_LINE] * (nv[S_OP0] - lv[S_OP0] - ONE));
        yc.constraint(lv[S_TSTORE] * (ONE - lv[IS_EXT_LINE]) * (lv[OP0] - nv[AUX0]));
        yc.constraint(lv[S_TLOAD] * (ONE - lv[IS_EXT_LINE]) * (lv[DST] - nv[AUX0]));
        yc.constraint(is_in_same_tx * (ONE - lv[S_TSTORE] - nv[S_CALL_SC]) * (nv[TP] - lv[TP]));
        yc.constraint(lv[S_TSTORE] * nv[IS_EXT_LINE] * (nv[TP] - lv[TP]));
        yc.constraint(lv[S_TSTORE] * (ONE - nv[IS_EXT_LINE]) * (nv[TP] - lv[S_OP0] - ONE));
        yc.constraint((ONE - lv[S_CALL_SC]) * nv[S_CALL_SC] * (nv[TP] - lv[TP]));
        yc.constraint(lv[S_CALL_SC] * (ONE - lv[IS_EXT_LINE]) * (nv[TP] - lv[TP]));
        yc.constraint(lv[S_CALL_SC] * lv[IS_EXT_LINE] * (nv[TP] - lv[TP] - T::c(12)));
        yc.constraint(lv[FILTER_TAPE_LOOKING] * (ONE - lv[FILTER_TAPE_LOOKING]));
        yc.constraint(lv[FILTER_TAPE_LOOKING] * (ONE - lv[S_TLOAD] - lv[S_TSTORE]));
        yc.constraint(lv[FILTER_TAPE_LOOKING] * (ONE - lv[IS_EXT_LINE]));
        yc.constraint((lv[S_TLOAD] + lv[S_TSTORE]) * lv[IS_EXT_LINE] * (ONE - lv[FILTER_TAPE_LOOKING]));
    }
    /* ---- call_sc.rs ---- */
    {
        for (int i = 0; i < 4; i++) yc.constraint(lv[S_CALL_SC] * (ONE - lv[IS_EXT_LINE]) * (nv[S_OP0 + i] - lv[ADDR_STORAGE + i]));
        for (int i = 0; i < 4; i++) yc.constraint(lv[S_CALL_SC] * (ONE - lv[IS_EXT_LINE]) * (nv[S_OP0 + 4 + i] - lv[ADDR_CODE + i]));
        yc.constraint(lv[S_CALL_SC] * (ONE - lv[IS_EXT_LINE]) * (nv[OP0] - lv[OP0]));
        yc.constraint(lv[S_CALL_SC] * (ONE - lv[IS_EXT_LINE]) * (nv[OP1] - lv[OP1]));
        yc.constraint_transition(lv[S_END] * (ONE - is_crossing_inst) * (lv[ENV_IDX] - nv[AUX0]));
        yc.constraint_transition(lv[S_END] * (ONE - is_crossing_inst) * (lv[CLK] - nv[AUX1]));
        yc.constraint(lv[S_CALL_SC] * is_crossing_inst * nv[CLK]);
        yc.constraint(lv[S_CALL_SC] * is_crossing_inst * nv[PC]);
        for (int i = 0; i < REGISTER_NUM; i++) yc.constraint(lv[S_CALL_SC] * is_crossing_inst * nv[REGS + i]);
        for (int i = 0; i < CTX_REGISTER_NUM; i++) {
            yc.constraint(lv[S_CALL_SC] * is_crossing_inst * (nv[ADDR_STORAGE + i] - lv[ADDR_STORAGE + i]));
            yc.constraint(lv[S_CALL_SC] * is_crossing_inst * (nv[ADDR_CODE + i] - lv[ADDR_CODE + i]));
        }
        yc.constraint(lv[S_END] * lv[IS_EXT_LINE] * (ONE - is_crossing_inst) * (nv[PC] - lv[PC]));
        yc.constraint(lv[S_END] * lv[IS_EXT_LINE] * (ONE - is_crossing_inst) * (nv[CLK] - lv[CLK]));
        yc.constraint(lv[IS_SCCALL_EXT_LINE] * (ONE - lv[IS_SCCALL_EXT_LINE]));
        yc.constraint((ONE - lv[S_CALL_SC]) * lv[IS_SCCALL_EXT_LINE]);
        yc.constraint(lv[S_CALL_SC] * lv[IS_EXT_LINE] * (ONE - lv[IS_SCCALL_EXT_LINE]));
        yc.constraint(lv[S_CALL_SC] * (ONE - lv[IS_EXT_LINE]) * lv[IS_SCCALL_EXT_LINE]);
        yc.constraint(lv[FILTER_SCCALL_END] * (ONE - lv[FILTER_SCCALL_END]));
        yc.constraint((ONE - lv[S_END]) * lv[FILTER_SCCALL_END]);
        yc.constraint(lv[S_END] * (ONE - lv[IS_EXT_LINE]) * lv[FILTER_SCCALL_END]);
        yc.constraint(lv[S_END] * lv[IS_EXT_LINE] * (ONE - lv[FILTER_SCCALL_END]));
    }
}

}  // namespace cpu_air
}  // namespace orc
#endif
