// AIR constraints of the CPU table (CpuStark, 94 columns, constraint degree 7), written once over an abstract
// field-value type T so that the same transcription is compiled (a) by nvcc into the device quotient kernel
// (T = air::Fp, one thread per LDE row) and (b) by g++ into the CPU oracle (T = base field for the quotient, T =
// quadratic extension for verify_proof).  NOTE: because the oracle includes this header, GPU-vs-
// oracle parity pins the evaluation machinery around these constraints, not the transcription itself; the
// transcription is checked by review against the cited Rust lines and by the "valid trace => quotient divisible /
// verifier accepts" tests.
//
// Constraints are emitted in the reference's source order (the consumer is Horner in alpha):
//   circuits/src/cpu/cpu_stark.rs:871-946  eval_packed_generic
//     :269-300  constraint_wrapper_cols        :302-339 constraint_tx_init      :675-714 constraint_ext_lines
//     :341-386  constraint_env_idx             :388-524 constraint_opcode_selector
//     :526-580  constraint_instruction_encode  :582-672 constraint_operands_mathches_registers
//     :716-741  constraint_env_unchanged_clk   :743-787 constraint_env_unchanged_pc
//     :789-826  constraint_reg_consistency     :828-866 CpuAdjacentRowWrapper::from_vars
//   circuits/src/cpu/{simple_arithmatic_op,mov,call,ret,mload,mstore,storage,tape,call_sc}.rs
//   column indices: circuits/src/cpu/columns.rs:4-133;  opcode masks: core/src/vm/opcodes.rs:76-105
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AIR_FN __host__ __device__ __forceinline__
#else
#define AIR_FN inline
#endif

namespace ola {
namespace air {

// canonical constant -> field value; specialised by each side for its value type
template <class T>
AIR_FN T kc(uint64_t k);

namespace cpu {
enum : int {
    COL_TX_IDX = 0, COL_ENV_IDX = 1, COL_CALL_SC_CNT = 2, COL_ADDR_STORAGE = 3, COL_ADDR_CODE = 7, COL_TP = 11, COL_CLK = 12, COL_PC = 13,
    COL_IS_EXT_LINE = 14, COL_EXT_CNT = 15, COL_REGS = 16, COL_INST = 26, COL_OP1_IMM = 27, COL_OPCODE = 28, COL_IMM_VAL = 29, COL_OP0 = 30,
    COL_OP1 = 31, COL_DST = 32, COL_AUX0 = 33, COL_AUX1 = 34, COL_IDX_STORAGE = 35, COL_S_OP0 = 36, COL_S_OP1 = 46, COL_S_DST = 56,
    COL_S_SIMPLE_ARITHMATIC_OP = 66, COL_S_MOV = 67, COL_S_JMP = 68, COL_S_CJMP = 69, COL_S_CALL = 70, COL_S_RET = 71, COL_S_MLOAD = 72,
    COL_S_MSTORE = 73, COL_S_END = 74, COL_S_RC = 75, COL_S_BITWISE = 76, COL_S_NOT = 77, COL_S_GTE = 78, COL_S_PSDN = 79, COL_S_SLOAD = 80,
    COL_S_SSTORE = 81, COL_S_TLOAD = 82, COL_S_TSTORE = 83, COL_S_CALL_SC = 84, NUM_OP_SELECTOR = 19, COL_IS_ENTRY_SC = 85,
    COL_IS_NEXT_LINE_DIFF_INST = 86, COL_IS_NEXT_LINE_SAME_TX = 87, COL_FILTER_TAPE_LOOKING = 88, IS_SCCALL_EXT_LINE = 89,
    COL_IS_STORAGE_EXT_LINE = 90, COL_FILTER_SCCALL_END = 91, COL_FILTER_LOOKING_PROG_IMM = 92, COL_IS_PADDING = 93, NUM_CPU_COLS = 94,
    REGISTER_NUM = 10, CTX_REGISTER_NUM = 4
};
// OlaOpcode::binary_bit_mask (core/src/vm/opcodes.rs:76-109)
static constexpr uint64_t OP_ADD = 1ull << 31, OP_MUL = 1ull << 30, OP_EQ = 1ull << 29, OP_ASSERT = 1ull << 28, OP_MOV = 1ull << 27,
                          OP_JMP = 1ull << 26, OP_CJMP = 1ull << 25, OP_CALL = 1ull << 24, OP_RET = 1ull << 23, OP_MLOAD = 1ull << 22,
                          OP_MSTORE = 1ull << 21, OP_END = 1ull << 20, OP_RC = 1ull << 19, OP_AND = 1ull << 18, OP_OR = 1ull << 17,
                          OP_XOR = 1ull << 16, OP_NOT = 1ull << 15, OP_NEQ = 1ull << 14, OP_GTE = 1ull << 13, OP_POSEIDON = 1ull << 12,
                          OP_SLOAD = 1ull << 11, OP_SSTORE = 1ull << 10, OP_TLOAD = 1ull << 9, OP_TSTORE = 1ull << 8, OP_SCCALL = 1ull << 7;

template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    const T one = kc<T>(1);
    // ---- CpuAdjacentRowWrapper::from_vars (cpu_stark.rs:828-866)
    const T lv_is_padding = lv[COL_IS_PADDING], nv_is_padding = nv[COL_IS_PADDING];
    const T lv_is_ext_inst = lv[COL_S_SLOAD] + lv[COL_S_SSTORE] + lv[COL_S_TLOAD] + lv[COL_S_TSTORE] + lv[COL_S_CALL_SC] + lv[COL_S_END];
    const T nv_is_ext_inst = nv[COL_S_SLOAD] + nv[COL_S_SSTORE] + nv[COL_S_TLOAD] + nv[COL_S_TSTORE] + nv[COL_S_CALL_SC] + nv[COL_S_END];
    const T lv_is_entry_sc = lv[COL_IS_ENTRY_SC];
    const T lv_ext_length = lv[COL_S_SLOAD] + lv[COL_S_SSTORE] + lv[COL_S_TLOAD] * (lv[COL_OP0] * lv[COL_OP1] + (one - lv[COL_OP0])) +
                            lv[COL_S_TSTORE] * lv[COL_OP1] + lv[COL_S_CALL_SC] + lv[COL_S_END] * (one - lv_is_entry_sc);
    const T is_crossing_inst = lv[COL_IS_NEXT_LINE_DIFF_INST];
    const T is_in_same_tx = lv[COL_IS_NEXT_LINE_SAME_TX];

    // ---- constraint_wrapper_cols (:269-300)
    yc.constraint(lv_is_padding * (lv_is_padding - one));
    yc.constraint_transition((nv_is_padding - lv_is_padding) * (nv_is_padding - lv_is_padding - one));
    yc.constraint(lv_is_padding * (lv[COL_S_END] - one));
    yc.constraint(lv_is_entry_sc * nv[COL_ENV_IDX]);
    yc.constraint((one - nv_is_padding) * is_in_same_tx * (nv[COL_TX_IDX] - lv[COL_TX_IDX]));
    yc.constraint_transition((one - nv_is_padding) * (one - is_in_same_tx) * (nv[COL_TX_IDX] - lv[COL_TX_IDX] - one));
    yc.constraint(is_crossing_inst * (lv_ext_length - lv[COL_EXT_CNT]));

    // ---- constraint_tx_init (:302-339)
    yc.constraint_first_row(lv[COL_TX_IDX]);
    yc.constraint_first_row(lv[COL_ENV_IDX]);
    yc.constraint_first_row(lv[COL_CALL_SC_CNT]);
    yc.constraint_first_row(lv[COL_CLK]);
    yc.constraint_first_row(lv[COL_PC]);
    for (int i = 0; i < REGISTER_NUM; ++i) yc.constraint_first_row(lv[COL_REGS + i]);
    yc.constraint_transition(is_in_same_tx * (nv[COL_TX_IDX] - lv[COL_TX_IDX]));
    yc.constraint_transition((one - is_in_same_tx) * nv[COL_ENV_IDX]);
    yc.constraint_transition((one - is_in_same_tx) * nv[COL_CALL_SC_CNT]);
    yc.constraint_transition((one - is_in_same_tx) * nv[COL_TP]);
    yc.constraint_transition((one - is_in_same_tx) * nv[COL_CLK]);
    yc.constraint_transition((one - is_in_same_tx) * nv[COL_PC]);
    for (int i = 0; i < REGISTER_NUM; ++i) yc.constraint_transition((one - is_in_same_tx) * nv[COL_REGS + i]);

    // ---- eval_packed_generic body (:884-929)
    yc.constraint_transition((one - nv_is_padding) * (one - lv[COL_S_END]) * (nv[COL_TX_IDX] - lv[COL_TX_IDX]));
    yc.constraint_transition((one - nv_is_padding) * lv_is_entry_sc * lv[COL_S_END] * (nv[COL_TX_IDX] - lv[COL_TX_IDX] - one));
    for (int i = 0; i < CTX_REGISTER_NUM; ++i) {
        yc.constraint_transition((one - nv_is_padding) * (one - lv[COL_S_END]) * (one - lv[COL_S_CALL_SC]) *
                                 (nv[COL_ADDR_STORAGE + i] - lv[COL_ADDR_STORAGE + i]));
        yc.constraint_transition((one - nv_is_padding) * (one - lv[COL_S_END]) * (one - lv[COL_S_CALL_SC]) *
                                 (nv[COL_ADDR_CODE + i] - lv[COL_ADDR_CODE + i]));
    }
    yc.constraint((one - lv[COL_IS_PADDING] - lv[COL_IS_EXT_LINE]) * lv[COL_OP1_IMM] * (one - lv[COL_FILTER_LOOKING_PROG_IMM]));
    yc.constraint((one - lv[COL_IS_PADDING] - lv[COL_IS_EXT_LINE]) * (lv[COL_S_MLOAD] + lv[COL_S_MSTORE]) * (one - lv[COL_FILTER_LOOKING_PROG_IMM]));

    // ---- constraint_ext_lines (:675-714)
    yc.constraint((one - lv_is_ext_inst) * lv[COL_IS_EXT_LINE]);
    yc.constraint(lv_is_ext_inst * (lv_ext_length - lv[COL_EXT_CNT]) * (one - nv[COL_IS_EXT_LINE]));
    yc.constraint(lv_is_ext_inst * (one - lv[COL_IS_EXT_LINE]) * lv[COL_EXT_CNT]);
    yc.constraint(nv_is_ext_inst * nv[COL_IS_EXT_LINE] * (nv[COL_EXT_CNT] - lv[COL_EXT_CNT] - one));
    yc.constraint(nv[COL_IS_EXT_LINE] * (nv[COL_OPCODE] - lv[COL_OPCODE]));
    for (int c = COL_S_SIMPLE_ARITHMATIC_OP; c < COL_S_SIMPLE_ARITHMATIC_OP + NUM_OP_SELECTOR; ++c)
        yc.constraint(nv[COL_IS_EXT_LINE] * (nv[c] - lv[c]));
    yc.constraint(nv[COL_IS_EXT_LINE] * (nv[COL_OP1_IMM] - lv[COL_OP1_IMM]));

    // ---- constraint_env_idx (:341-386)
    yc.constraint_transition(lv[COL_S_CALL_SC] * is_crossing_inst * (nv[COL_CALL_SC_CNT] - lv[COL_CALL_SC_CNT] - one));
    yc.constraint_transition(is_in_same_tx * (one - lv[COL_S_CALL_SC]) * (nv[COL_CALL_SC_CNT] - lv[COL_CALL_SC_CNT]));
    yc.constraint(lv[COL_S_CALL_SC] * (one - is_crossing_inst) * (nv[COL_CALL_SC_CNT] - lv[COL_CALL_SC_CNT]));
    yc.constraint(lv[COL_S_CALL_SC] * is_crossing_inst * (nv[COL_ENV_IDX] - lv[COL_CALL_SC_CNT]));
    yc.constraint((one - lv[COL_S_CALL_SC] - lv[COL_S_END]) * (nv[COL_ENV_IDX] - lv[COL_ENV_IDX]));
    yc.constraint(lv[COL_S_CALL_SC] * (one - is_crossing_inst) * (nv[COL_ENV_IDX] - lv[COL_ENV_IDX]));
    yc.constraint(lv[COL_S_END] * lv[COL_IS_EXT_LINE] * (nv[COL_ENV_IDX] - lv[COL_ENV_IDX]));

    // ---- constraint_opcode_selector (:388-524)
    {
        const int sel[NUM_OP_SELECTOR] = {COL_S_SIMPLE_ARITHMATIC_OP, COL_S_MOV, COL_S_JMP, COL_S_CJMP, COL_S_CALL, COL_S_RET, COL_S_MLOAD,
                                          COL_S_MSTORE, COL_S_END, COL_S_RC, COL_S_BITWISE, COL_S_NOT, COL_S_GTE, COL_S_PSDN, COL_S_SLOAD,
                                          COL_S_SSTORE, COL_S_TLOAD, COL_S_TSTORE, COL_S_CALL_SC};
        const uint64_t opc[NUM_OP_SELECTOR] = {0, OP_MOV, OP_JMP, OP_CJMP, OP_CALL, OP_RET, OP_MLOAD, OP_MSTORE, OP_END, OP_RC, 0, OP_NOT,
                                               OP_GTE, OP_POSEIDON, OP_SLOAD, OP_SSTORE, OP_TLOAD, OP_TSTORE, OP_SCCALL};
        const T opcode = lv[COL_OPCODE];
        yc.constraint(lv[COL_S_SIMPLE_ARITHMATIC_OP] * (opcode - kc<T>(OP_ADD)) * (opcode - kc<T>(OP_MUL)) * (opcode - kc<T>(OP_EQ)) *
                      (opcode - kc<T>(OP_NEQ)) * (opcode - kc<T>(OP_ASSERT)));
        yc.constraint(lv[COL_S_BITWISE] * (opcode - kc<T>(OP_AND)) * (opcode - kc<T>(OP_OR)) * (opcode - kc<T>(OP_XOR)));
        for (int i = 0; i < NUM_OP_SELECTOR; ++i) {
            const T s = lv[sel[i]];
            yc.constraint(s * (one - s));
        }
        T sum_s_op = kc<T>(0);
        for (int i = 0; i < NUM_OP_SELECTOR; ++i) sum_s_op = sum_s_op + lv[sel[i]];
        yc.constraint(one - sum_s_op);
        T cal_opcode = kc<T>(0);
        for (int i = 0; i < NUM_OP_SELECTOR; ++i) cal_opcode = cal_opcode + lv[sel[i]] * kc<T>(opc[i]);
        yc.constraint((opcode - cal_opcode) * (one - lv[COL_S_BITWISE] - lv[COL_S_SIMPLE_ARITHMATIC_OP]));
    }

    // ---- constraint_instruction_encode (:526-580): OP1_IMM_SHIFT 62, OP0/OP1/DST shift starts 61/51/41, registers r9..r0
    {
        yc.constraint(lv[COL_OP1_IMM] * (one - lv[COL_OP1_IMM]));
        T instruction = lv[COL_OP1_IMM] * kc<T>(1ull << 62);
        for (int index = 0; index < REGISTER_NUM; ++index) instruction = instruction + lv[COL_S_OP0 + REGISTER_NUM - 1 - index] * kc<T>((1ull << 61) >> index);
        for (int index = 0; index < REGISTER_NUM; ++index) instruction = instruction + lv[COL_S_OP1 + REGISTER_NUM - 1 - index] * kc<T>((1ull << 51) >> index);
        for (int index = 0; index < REGISTER_NUM; ++index) instruction = instruction + lv[COL_S_DST + REGISTER_NUM - 1 - index] * kc<T>((1ull << 41) >> index);
        instruction = instruction + lv[COL_OPCODE];
        yc.constraint((one - lv[COL_IS_EXT_LINE]) * (lv[COL_INST] - instruction));
        yc.constraint((one - lv[COL_IS_EXT_LINE]) * (lv[COL_OP1_IMM] * (lv[COL_OP1] - lv[COL_IMM_VAL])));
    }

    // ---- constraint_operands_mathches_registers (:582-672)
    {
        const T not_ext = one - lv[COL_IS_EXT_LINE];
        for (int i = 0; i < REGISTER_NUM; ++i) { const T s = lv[COL_S_OP0 + i]; yc.constraint(not_ext * s * (one - s)); }
        for (int i = 0; i < REGISTER_NUM; ++i) { const T s = lv[COL_S_OP1 + i]; yc.constraint(not_ext * s * (one - s)); }
        for (int i = 0; i < REGISTER_NUM; ++i) { const T s = lv[COL_S_DST + i]; yc.constraint(not_ext * s * (one - s)); }
        T sum_s_op0 = kc<T>(0), sum_s_op1 = kc<T>(0), sum_s_dst = kc<T>(0);
        for (int i = 0; i < REGISTER_NUM; ++i) sum_s_op0 = sum_s_op0 + lv[COL_S_OP0 + i];
        yc.constraint(not_ext * sum_s_op0 * (one - sum_s_op0));
        for (int i = 0; i < REGISTER_NUM; ++i) sum_s_op1 = sum_s_op1 + lv[COL_S_OP1 + i];
        yc.constraint(not_ext * sum_s_op1 * (one - sum_s_op1));
        for (int i = 0; i < REGISTER_NUM; ++i) sum_s_dst = sum_s_dst + lv[COL_S_DST + i];
        yc.constraint(not_ext * sum_s_dst * (one - sum_s_dst));
        T op0_sum = kc<T>(0), op1_sum = kc<T>(0), dst_sum = kc<T>(0);
        for (int i = 0; i < REGISTER_NUM; ++i) op0_sum = op0_sum + lv[COL_S_OP0 + i] * lv[COL_REGS + i];
        yc.constraint(not_ext * sum_s_op0 * (lv[COL_OP0] - op0_sum));
        for (int i = 0; i < REGISTER_NUM; ++i) op1_sum = op1_sum + lv[COL_S_OP1 + i] * lv[COL_REGS + i];
        yc.constraint(not_ext * sum_s_op1 * (lv[COL_OP1] - op1_sum));
        for (int i = 0; i < REGISTER_NUM; ++i) dst_sum = dst_sum + lv[COL_S_DST + i] * nv[COL_REGS + i];
        yc.constraint(not_ext * sum_s_dst * (lv[COL_DST] - dst_sum));
    }

    // ---- constraint_env_unchanged_clk (:716-741)
    yc.constraint(nv[COL_IS_EXT_LINE] * (one - nv[COL_S_END]) * (nv[COL_CLK] - lv[COL_CLK]));
    yc.constraint(is_in_same_tx * (one - lv[COL_S_CALL_SC] - lv[COL_S_END]) * (one - nv[COL_IS_EXT_LINE]) * (nv[COL_CLK] - lv[COL_CLK] - one));

    // ---- constraint_env_unchanged_pc (:743-787)  (its first constraint repeats the clk one verbatim in the reference)
    {
        yc.constraint(nv[COL_IS_EXT_LINE] * (one - nv[COL_S_END]) * (nv[COL_CLK] - lv[COL_CLK]));
        const T instruction_size = (one - lv[COL_S_MLOAD] - lv[COL_S_MSTORE]) * (one + lv[COL_OP1_IMM]) + (lv[COL_S_MLOAD] + lv[COL_S_MSTORE]) * kc<T>(2);
        const T pc_incr = (one - (lv[COL_S_JMP] + lv[COL_S_CJMP] + lv[COL_S_CALL] + lv[COL_S_RET])) * (lv[COL_PC] + instruction_size);
        const T pc_jmp = lv[COL_S_JMP] * lv[COL_OP1];
        const T pc_cjmp = lv[COL_S_CJMP] * ((one - lv[COL_OP0]) * (lv[COL_PC] + instruction_size) + lv[COL_OP0] * lv[COL_OP1]);
        const T pc_call = lv[COL_S_CALL] * lv[COL_OP1];
        const T pc_ret = lv[COL_S_RET] * lv[COL_DST];
        yc.constraint((one - nv[COL_IS_EXT_LINE]) * (one - lv[COL_S_END] - lv[COL_S_CALL_SC]) *
                      (nv[COL_PC] - (pc_incr + pc_jmp + pc_cjmp + pc_call + pc_ret)));
        yc.constraint((one - nv[COL_IS_EXT_LINE]) * lv[COL_S_CJMP] * lv[COL_OP0] * (one - lv[COL_OP0]));
    }

    // ---- constraint_reg_consistency (:789-826)
    {
        const T multi_reg_change = lv[COL_S_SLOAD] + lv[COL_S_PSDN] + lv[COL_S_CALL_SC] * is_crossing_inst + lv[COL_S_END] * (one - lv[COL_IS_EXT_LINE]);
        for (int i = 0; i < REGISTER_NUM - 1; ++i)
            yc.constraint_transition((one - multi_reg_change) * (one - lv[COL_S_DST + i]) * (nv[COL_REGS + i] - lv[COL_REGS + i]));
        yc.constraint_transition((one - lv[COL_S_RET] - lv[COL_S_CALL_SC] * is_crossing_inst - lv[COL_S_END]) * (one - lv[COL_S_DST + REGISTER_NUM - 1]) *
                                 (nv[COL_REGS + REGISTER_NUM - 1] - lv[COL_REGS + REGISTER_NUM - 1]));
    }

    // ---- simple_arithmatic_op.rs
    {
        const T s = lv[COL_S_SIMPLE_ARITHMATIC_OP], opcode = lv[COL_OPCODE];
        const T d_add = opcode - kc<T>(OP_ADD), d_mul = opcode - kc<T>(OP_MUL), d_eq = opcode - kc<T>(OP_EQ), d_neq = opcode - kc<T>(OP_NEQ),
                d_assert = opcode - kc<T>(OP_ASSERT);
        const T is_add = s * d_mul * d_eq * d_neq * d_assert;
        const T is_mul = s * d_add * d_eq * d_neq * d_assert;
        const T is_eq = s * d_add * d_mul * d_neq * d_assert;
        const T is_neq = s * d_add * d_mul * d_eq * d_assert;
        const T is_assert = s * d_add * d_mul * d_eq * d_neq;
        yc.constraint(is_add * (lv[COL_DST] - (lv[COL_OP0] + lv[COL_OP1])));
        yc.constraint(is_mul * (lv[COL_DST] - lv[COL_OP0] * lv[COL_OP1]));
        const T op_diff = lv[COL_OP0] - lv[COL_OP1];
        const T diff_aux = op_diff * lv[COL_AUX0];
        const T res = lv[COL_DST];
        const T eq_cs = is_eq * (res * op_diff + (one - res) * (one - diff_aux));
        const T neq_cs = is_neq * ((one - res) * op_diff + res * (one - diff_aux));
        yc.constraint(eq_cs + neq_cs);
        yc.constraint(is_assert * (one - lv[COL_OP1]));
    }
    // ---- mov.rs
    yc.constraint(lv[COL_S_MOV] * (lv[COL_DST] - lv[COL_OP1]));
    // ---- call.rs
    {
        const T two = one + one;
        const T fp = lv[COL_REGS + REGISTER_NUM - 1];
        const T op0_cs = lv[COL_OP0] + one - fp;
        const T op1_cs = lv[COL_OP1_IMM] * (lv[COL_DST] - lv[COL_PC] - two) + (one - lv[COL_OP1_IMM]) * (lv[COL_DST] - lv[COL_PC] - one);
        const T aux0_cs = lv[COL_AUX0] - fp + two;
        yc.constraint(lv[COL_S_CALL] * (op0_cs + op1_cs + aux0_cs));
    }
    // ---- ret.rs
    {
        const T fp = lv[COL_REGS + REGISTER_NUM - 1];
        const T op0_cs = lv[COL_OP0] + one - fp;
        const T dst_cs = lv[COL_DST] - nv[COL_PC];
        const T aux0_cs = lv[COL_AUX0] + one + one - fp;
        yc.constraint(lv[COL_S_RET] * (op0_cs + dst_cs + aux0_cs));
        yc.constraint_transition(lv[COL_S_RET] * (nv[COL_REGS + REGISTER_NUM - 1] - lv[COL_AUX1]));
    }
    // ---- mload.rs / mstore.rs
    for (int k = 0; k < 2; ++k) {
        const T s = lv[k == 0 ? COL_S_MLOAD : COL_S_MSTORE];
        yc.constraint(s * (one - lv[COL_OP1_IMM]) * (lv[COL_AUX0] - lv[COL_IMM_VAL]));
        yc.constraint(s * lv[COL_OP1_IMM] * (lv[COL_AUX1] - lv[COL_OP0] - lv[COL_OP1]));
        yc.constraint(s * (one - lv[COL_OP1_IMM]) * (lv[COL_AUX1] - lv[COL_OP0] - lv[COL_AUX0] * lv[COL_OP1]));
    }
    // ---- storage.rs
    {
        const T st = lv[COL_S_SSTORE] + lv[COL_S_SLOAD];
        yc.constraint_first_row(lv[COL_IDX_STORAGE] - st);
        yc.constraint_transition(nv[COL_IDX_STORAGE] - lv[COL_IDX_STORAGE] - nv[COL_IS_STORAGE_EXT_LINE]);
        yc.constraint(st * (one - lv[COL_IS_EXT_LINE]) * (nv[COL_OP0] - lv[COL_OP0]));
        yc.constraint(st * (one - lv[COL_IS_EXT_LINE]) * (nv[COL_OP1] - lv[COL_OP1]));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP0] - lv[COL_OP0]));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP0 + 1] - lv[COL_S_OP0] - one));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP0 + 2] - lv[COL_S_OP0 + 1] - one));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP0 + 3] - lv[COL_S_OP0 + 2] - one));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP1] - lv[COL_OP1]));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP1 + 1] - lv[COL_S_OP1] - one));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP1 + 2] - lv[COL_S_OP1 + 1] - one));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (lv[COL_S_OP1 + 3] - lv[COL_S_OP1 + 2] - one));
        yc.constraint(st * lv[COL_IS_EXT_LINE] * (one - lv[COL_IS_STORAGE_EXT_LINE]));
        yc.constraint((one - st) * lv[COL_IS_STORAGE_EXT_LINE]);
        yc.constraint(st * (one - lv[COL_IS_EXT_LINE]) * lv[COL_IS_STORAGE_EXT_LINE]);
    }
    // ---- tape.rs
    {
        const T l_ts = lv[COL_S_TSTORE], l_tl = lv[COL_S_TLOAD], l_ext = lv[COL_IS_EXT_LINE], n_ext = nv[COL_IS_EXT_LINE];
        yc.constraint((nv[COL_S_TSTORE] + nv[COL_S_TLOAD]) * n_ext * (nv[COL_OP0] - lv[COL_OP0]));
        yc.constraint((nv[COL_S_TSTORE] + nv[COL_S_TLOAD]) * n_ext * (nv[COL_OP1] - lv[COL_OP1]));
        yc.constraint((l_ts + l_tl) * l_ext * n_ext * (nv[COL_AUX0] - lv[COL_AUX0] - one));
        yc.constraint(l_ts * (one - l_ext) * (lv[COL_TP] - nv[COL_S_OP0]));
        yc.constraint(l_ts * l_ext * n_ext * (nv[COL_S_OP0] - lv[COL_S_OP0] - one));
        yc.constraint(l_ts * (one - n_ext) * (nv[COL_TP] - lv[COL_S_OP0] - one));
        yc.constraint(l_tl * lv[COL_OP0] * (one - l_ext) * (nv[COL_S_OP0] + lv[COL_OP1] - lv[COL_TP]));
        yc.constraint(l_tl * (one - lv[COL_OP0]) * (one - l_ext) * (nv[COL_S_OP0] - lv[COL_OP1]));
        yc.constraint((l_ts + l_tl) * l_ext * n_ext * (nv[COL_S_OP0] - lv[COL_S_OP0] - one));
        yc.constraint(l_ts * (one - l_ext) * (lv[COL_OP0] - nv[COL_AUX0]));
        yc.constraint(l_tl * (one - l_ext) * (lv[COL_DST] - nv[COL_AUX0]));
        yc.constraint(is_in_same_tx * (one - l_ts - nv[COL_S_CALL_SC]) * (nv[COL_TP] - lv[COL_TP]));
        yc.constraint(l_ts * n_ext * (nv[COL_TP] - lv[COL_TP]));
        yc.constraint(l_ts * (one - n_ext) * (nv[COL_TP] - lv[COL_S_OP0] - one));
        yc.constraint((one - lv[COL_S_CALL_SC]) * nv[COL_S_CALL_SC] * (nv[COL_TP] - lv[COL_TP]));
        yc.constraint(lv[COL_S_CALL_SC] * (one - l_ext) * (nv[COL_TP] - lv[COL_TP]));
        yc.constraint(lv[COL_S_CALL_SC] * l_ext * (nv[COL_TP] - lv[COL_TP] - kc<T>(12)));
        yc.constraint(lv[COL_FILTER_TAPE_LOOKING] * (one - lv[COL_FILTER_TAPE_LOOKING]));
        yc.constraint(lv[COL_FILTER_TAPE_LOOKING] * (one - l_tl - l_ts));
        yc.constraint(lv[COL_FILTER_TAPE_LOOKING] * (one - l_ext));
        yc.constraint((l_tl + l_ts) * l_ext * (one - lv[COL_FILTER_TAPE_LOOKING]));
    }
    // ---- call_sc.rs
    {
        const T sc = lv[COL_S_CALL_SC], l_ext = lv[COL_IS_EXT_LINE], end = lv[COL_S_END];
        for (int i = 0; i < 4; ++i) yc.constraint(sc * (one - l_ext) * (nv[COL_S_OP0 + i] - lv[COL_ADDR_STORAGE + i]));
        for (int i = 0; i < 4; ++i) yc.constraint(sc * (one - l_ext) * (nv[COL_S_OP0 + 4 + i] - lv[COL_ADDR_CODE + i]));
        yc.constraint(sc * (one - l_ext) * (nv[COL_OP0] - lv[COL_OP0]));
        yc.constraint(sc * (one - l_ext) * (nv[COL_OP1] - lv[COL_OP1]));
        yc.constraint_transition(end * (one - is_crossing_inst) * (lv[COL_ENV_IDX] - nv[COL_AUX0]));
        yc.constraint_transition(end * (one - is_crossing_inst) * (lv[COL_CLK] - nv[COL_AUX1]));
        yc.constraint(sc * is_crossing_inst * nv[COL_CLK]);
        yc.constraint(sc * is_crossing_inst * nv[COL_PC]);
        for (int i = 0; i < REGISTER_NUM; ++i) yc.constraint(sc * is_crossing_inst * nv[COL_REGS + i]);
        for (int i = 0; i < CTX_REGISTER_NUM; ++i) {
            yc.constraint(sc * is_crossing_inst * (nv[COL_ADDR_STORAGE + i] - lv[COL_ADDR_STORAGE + i]));
            yc.constraint(sc * is_crossing_inst * (nv[COL_ADDR_CODE + i] - lv[COL_ADDR_CODE + i]));
        }
        yc.constraint(end * l_ext * (one - is_crossing_inst) * (nv[COL_PC] - lv[COL_PC]));
        yc.constraint(end * l_ext * (one - is_crossing_inst) * (nv[COL_CLK] - lv[COL_CLK]));
        yc.constraint(lv[IS_SCCALL_EXT_LINE] * (one - lv[IS_SCCALL_EXT_LINE]));
        yc.constraint((one - sc) * lv[IS_SCCALL_EXT_LINE]);
        yc.constraint(sc * l_ext * (one - lv[IS_SCCALL_EXT_LINE]));
        yc.constraint(sc * (one - l_ext) * lv[IS_SCCALL_EXT_LINE]);
        yc.constraint(lv[COL_FILTER_SCCALL_END] * (one - lv[COL_FILTER_SCCALL_END]));
        yc.constraint((one - end) * lv[COL_FILTER_SCCALL_END]);
        yc.constraint(end * (one - l_ext) * lv[COL_FILTER_SCCALL_END]);
        yc.constraint(end * l_ext * (one - lv[COL_FILTER_SCCALL_END]));
    }
}

}  // namespace cpu
}  // namespace air
}  // namespace ola
