/* ORACLE (test infrastructure, NOT product code): generate_cpu_trace restated.
 *
 *   orc_generate_cpu_trace   circuits/src/generation/cpu.rs:11-218
 *                            columns: circuits/src/cpu/columns.rs (94 columns); opcode masks: core/src/vm/opcodes.rs
 *                            (binary_bit_mask = 1 << binary_bit_shift; ADD 31 ... SCCALL 7)
 *
 * The executor's Step (core/src/trace/trace.rs) enters as one record of 66 u64 per executed row:
 *    0 env_idx   1 call_sc_cnt   2..5 addr_storage[4]   6..9 addr_code[4]   10 tp   11 clk   12 pc   13 is_ext_line   14 ext_cnt
 *   15..24 regs[10]   25 instruction   26 op1_imm   27 opcode   28 immediate_data
 *   29 op0   30 op1   31 dst   32 aux0   33 aux1 (register_selector)   34 storage_access_idx
 *   35..44 op0_reg_sel[10]   45..54 op1_reg_sel[10]   55..64 dst_reg_sel[10]   65 filter_tape_looking */
#include <string.h>

#include "oracle.h"

enum { OP_ADD = 31, OP_MUL = 30, OP_EQ = 29, OP_ASSERT = 28, OP_MOV = 27, OP_JMP = 26, OP_CJMP = 25, OP_CALL = 24, OP_RET = 23, OP_MLOAD = 22,
       OP_MSTORE = 21, OP_END = 20, OP_RC = 19, OP_AND = 18, OP_OR = 17, OP_XOR = 16, OP_NOT = 15, OP_NEQ = 14, OP_GTE = 13, OP_POSEIDON = 12,
       OP_SLOAD = 11, OP_SSTORE = 10, OP_TLOAD = 9, OP_TSTORE = 8, OP_SCCALL = 7 };
#define MASK(op) (1ull << (op))

/* opcode_to_selector (cpu.rs:19-59) */
static int selector_of(uint64_t opcode) {
    if (opcode == MASK(OP_ADD) || opcode == MASK(OP_MUL) || opcode == MASK(OP_EQ) || opcode == MASK(OP_ASSERT) || opcode == MASK(OP_NEQ)) return 66;
    if (opcode == MASK(OP_MOV)) return 67;
    if (opcode == MASK(OP_JMP)) return 68;
    if (opcode == MASK(OP_CJMP)) return 69;
    if (opcode == MASK(OP_CALL)) return 70;
    if (opcode == MASK(OP_RET)) return 71;
    if (opcode == MASK(OP_MLOAD)) return 72;
    if (opcode == MASK(OP_MSTORE)) return 73;
    if (opcode == MASK(OP_END)) return 74;
    if (opcode == MASK(OP_RC)) return 75;
    if (opcode == MASK(OP_AND) || opcode == MASK(OP_OR) || opcode == MASK(OP_XOR)) return 76;
    if (opcode == MASK(OP_NOT)) return 77;
    if (opcode == MASK(OP_GTE)) return 78;
    if (opcode == MASK(OP_POSEIDON)) return 79;
    if (opcode == MASK(OP_SLOAD)) return 80;
    if (opcode == MASK(OP_SSTORE)) return 81;
    if (opcode == MASK(OP_TLOAD)) return 82;
    if (opcode == MASK(OP_TSTORE)) return 83;
    if (opcode == MASK(OP_SCCALL)) return 84;
    return -1;
}

/* steps [nrows][66]; out [94][n] column-major, n a power of two >= max(nrows, 1) (the reference takes the next power of two). */
void orc_generate_cpu_trace(const uint64_t *steps, size_t nrows, size_t n, uint64_t *out) {
    memset(out, 0, 94 * n * sizeof(uint64_t));
#define T(c, i) out[(size_t)(c) * n + (i)]
    for (size_t i = 0; i < nrows; ++i) { /* :61-178 */
        const uint64_t *s = steps + i * 66;
        T(0, i) = 0;                     /* COL_TX_IDX */
        T(1, i) = gl_canon(s[0]);        /* COL_ENV_IDX */
        T(2, i) = gl_canon(s[1]);        /* COL_CALL_SC_CNT */
        for (int j = 0; j < 4; ++j) T(3 + j, i) = gl_canon(s[2 + j]), T(7 + j, i) = gl_canon(s[6 + j]);
        T(11, i) = gl_canon(s[10]);      /* TP */
        T(12, i) = (uint32_t)s[11];      /* CLK: from_canonical_u32 */
        T(13, i) = gl_canon(s[12]);      /* PC */
        T(14, i) = gl_canon(s[13]);      /* IS_EXT_LINE */
        T(15, i) = gl_canon(s[14]);      /* EXT_CNT */
        for (int j = 0; j < 10; ++j) T(16 + j, i) = gl_canon(s[15 + j]);
        T(26, i) = gl_canon(s[25]), T(27, i) = gl_canon(s[26]), T(28, i) = gl_canon(s[27]), T(29, i) = gl_canon(s[28]);
        T(30, i) = gl_canon(s[29]), T(31, i) = gl_canon(s[30]), T(32, i) = gl_canon(s[31]), T(33, i) = gl_canon(s[32]), T(34, i) = gl_canon(s[33]);
        T(35, i) = gl_canon(s[34]);      /* IDX_STORAGE */
        for (int j = 0; j < 10; ++j) T(36 + j, i) = gl_canon(s[35 + j]), T(46 + j, i) = gl_canon(s[45 + j]), T(56 + j, i) = gl_canon(s[55 + j]);
        const uint64_t opcode = s[27], env = s[0], op0 = s[29], op1 = s[30], ext_cnt = s[14], is_ext = s[13];
        const int sel = selector_of(opcode);
        if (sel >= 0) T(sel, i) = 1;
        const int env_zero = gl_canon(env) == 0;
        T(85, i) = env_zero ? 1 : 0; /* IS_ENTRY_SC */
        uint64_t ext_length; /* :118-132, plain u64 arithmetic on the inner values */
        if (opcode == MASK(OP_SLOAD) || opcode == MASK(OP_SSTORE) || opcode == MASK(OP_SCCALL) || (opcode == MASK(OP_END) && !env_zero))
            ext_length = 1;
        else if (opcode == MASK(OP_TLOAD))
            ext_length = op0 * op1 + (1 - op0);
        else if (opcode == MASK(OP_TSTORE))
            ext_length = op1;
        else
            ext_length = 0;
        T(86, i) = ext_length == ext_cnt ? 1 : 0;                              /* IS_NEXT_LINE_DIFF_INST */
        T(87, i) = (env_zero && opcode == MASK(OP_END)) ? 0 : 1;               /* IS_NEXT_LINE_SAME_TX */
        T(88, i) = gl_canon(s[65]);                                            /* FILTER_TAPE_LOOKING */
        T(89, i) = (opcode == MASK(OP_SCCALL) && ext_cnt == 1) ? 1 : 0;        /* IS_SCCALL_EXT_LINE */
        T(90, i) = ((opcode == MASK(OP_SLOAD) || opcode == MASK(OP_SSTORE)) && is_ext == 1) ? 1 : 0; /* IS_STORAGE_EXT_LINE */
        T(91, i) = (opcode == MASK(OP_END) && is_ext == 1) ? 1 : 0;            /* FILTER_SCCALL_END */
        T(92, i) = is_ext == 1 ? 0 : ((opcode == MASK(OP_MLOAD) || opcode == MASK(OP_MSTORE)) ? 1 : (s[26] == 1 ? 1 : 0)); /* FILTER_LOOKING_PROG_IMM */
    }
    /* padding (:180-208) */
    const uint64_t inst_end = nrows == 0 ? 1048576 : T(26, nrows - 1);
    const uint64_t last_tx = nrows == 0 ? 0 : T(0, nrows - 1);
    const uint64_t last_idx_storage = nrows == 0 ? 0 : T(35, nrows - 1);
    for (size_t i = nrows; i < n; ++i) {
        T(0, i) = last_tx, T(26, i) = inst_end, T(28, i) = MASK(OP_END), T(35, i) = last_idx_storage;
        T(74, i) = 1, T(85, i) = 1, T(86, i) = 1, T(87, i) = 0, T(93, i) = 1;
    }
#undef T
}

/* ---- generate_memory_trace (circuits/src/generation/memory.rs:8-155) -----------------------------------------------------------
 * The executor's MemoryTraceCell list (core/src/trace/trace.rs; already sorted and differenced by gen_memory_table) enters as
 * one record of 15 u64 per cell:
 *    0 env_idx   1 is_rw   2 addr   3 clk   4 op   5 is_write   6 value   7 diff_addr   8 diff_addr_inv   9 diff_clk
 *   10 diff_addr_cond   11 rw_addr_unchanged   12 region_prophet   13 region_heap   14 rc_value
 * Columns: circuits/src/memory/columns.rs:12-45 (29 columns).  out [29][n], n a power of two >= max(ncells, 2). */
static int mem_selector_of(uint64_t op) {
    if (op == 0) return 16;                 /* COL_MEM_S_PROPHET */
    if (op == MASK(OP_MLOAD)) return 6;
    if (op == MASK(OP_MSTORE)) return 7;
    if (op == MASK(OP_CALL)) return 8;
    if (op == MASK(OP_RET)) return 9;
    if (op == MASK(OP_TLOAD)) return 10;
    if (op == MASK(OP_TSTORE)) return 11;
    if (op == MASK(OP_SCCALL)) return 12;
    if (op == MASK(OP_POSEIDON)) return 13;
    if (op == MASK(OP_SSTORE)) return 14;
    if (op == MASK(OP_SLOAD)) return 15;
    return -1;
}
void orc_generate_memory_trace(const uint64_t *cells, size_t ncells, size_t n, uint64_t *out) {
    const uint64_t SPAN = 0xFFFFFFFFull; /* 2^32 - 1 */
    memset(out, 0, 29 * n * sizeof(uint64_t));
#define T(c, i) out[(size_t)(c) * n + (i)]
    size_t filled = ncells;
    for (size_t i = 0; i < ncells; ++i) { /* :51-99 */
        const uint64_t *c = cells + i * 15;
        T(0, i) = 0;
        T(1, i) = gl_canon(c[0]), T(2, i) = gl_canon(c[1]), T(3, i) = gl_canon(c[2]), T(4, i) = gl_canon(c[3]), T(5, i) = gl_canon(c[4]);
        const int sel = mem_selector_of(c[4]);
        if (sel >= 0) T(sel, i) = 1;
        T(17, i) = gl_canon(c[5]), T(18, i) = gl_canon(c[6]), T(19, i) = gl_canon(c[7]), T(20, i) = gl_canon(c[8]), T(21, i) = gl_canon(c[9]);
        T(22, i) = gl_canon(c[10]), T(23, i) = gl_canon(c[11]), T(24, i) = gl_canon(c[12]), T(25, i) = gl_canon(c[13]), T(26, i) = gl_canon(c[14]);
        const int prophet = gl_canon(c[12]) == 1, heap = gl_canon(c[13]) == 1;
        const int last_is_not_heap = i > 0 && gl_canon(cells[(i - 1) * 15 + 13]) == 0;
        T(27, i) = (i == 0 || prophet || (heap && last_is_not_heap)) ? 0 : 1;   /* FILTER_LOOKING_RC */
        T(28, i) = (heap || prophet) ? 1 : 0;                                    /* FILTER_LOOKING_RC_COND */
    }
    if (filled == 0) { /* :101-112 */
        const uint64_t addr = gl_sub(0, SPAN);
        T(3, 0) = addr, T(17, 0) = 1, T(22, 0) = gl_sub(0, addr), T(24, 0) = 1, T(26, 0) = gl_sub(0, addr);
        filled = 1;
    }
    if (n != filled) { /* :115-146 */
        uint64_t addr = T(2, filled - 1) == 1 ? gl_sub(0, SPAN) : gl_add(T(3, filled - 1), 1);
        const uint64_t tx_idx = T(0, filled - 1), env_idx = T(1, filled - 1);
        int first = 1;
        for (size_t i = filled; i < n; ++i) {
            T(16, i) = 1, T(0, i) = tx_idx, T(1, i) = env_idx, T(3, i) = addr, T(17, i) = 1;
            T(19, i) = first ? gl_sub(addr, T(3, filled - 1)) : 1;
            T(20, i) = gl_inv(T(19, i));
            T(22, i) = gl_sub(0, addr);
            T(24, i) = 1;
            T(26, i) = T(22, i);
            addr = gl_add(addr, 1);
            first = 0;
        }
    }
#undef T
}

/* ---- generate_prog_trace (circuits/src/generation/prog.rs:18-157) -----------------------------------------------------------------
 * steps: the 66-u64 Step records of orc_generate_cpu_trace (is_ext_line 13, op1_imm 26, opcode 27, addr_code 6..9, pc 12,
 * instruction 25, immediate_data 28); prog_rows [m][6] = (addr0..3, pc, inst), the lines of every program in order (the Rust
 * flattens `progs` in the same order, prog.rs:106-131); roots[8] = start_root[4], end_root[4].  Columns program/columns.rs:3-16
 * (18).  Returns the number of rows n = max(next_power_of_two(max(exec_len, m)), 2) when out == NULL or too small. */
uint64_t orc_compress_challenge(const uint64_t *const *cols, uint32_t ncols, size_t n); /* stark_api.cpp */
static uint64_t prog_compress(const uint64_t v[6], uint64_t beta) { /* prog.rs:149-156 */
    uint64_t acc = 0;
    for (int k = 5; k >= 0; --k) acc = gl_add(gl_mul(acc, beta), gl_canon(v[k]));
    return acc;
}
size_t orc_generate_prog_trace(const uint64_t *steps, size_t nsteps, const uint64_t *prog_rows, size_t m, const uint64_t roots[8], uint64_t *out,
                               size_t out_cap_rows, uint64_t *beta_out) {
    uint64_t inter[8]; /* observe start[i], end[i] for i in 0..4 (prog.rs:25-28) */
    for (int i = 0; i < 4; ++i) inter[2 * i] = roots[i], inter[2 * i + 1] = roots[4 + i];
    const uint64_t *col = inter;
    const uint64_t beta = orc_compress_challenge(&col, 1, 8);
    size_t exec_len = 0;
    for (size_t i = 0; i < nsteps; ++i) {
        const uint64_t *s = steps + i * 66;
        if (s[13] == 1) continue;
        exec_len += (s[26] == 1 || s[27] == MASK(OP_MLOAD) || s[27] == MASK(OP_MSTORE)) ? 2 : 1;
    }
    size_t filled = exec_len > m ? exec_len : m, n = 2;
    while (n < filled) n <<= 1;
    if (out == NULL || out_cap_rows < n) return n;
    n = out_cap_rows; /* a caller may ask for a larger power of two */
    memset(out, 0, 18 * n * sizeof(uint64_t));
#define T(c, i) out[(size_t)(c) * n + (i)]
    size_t e = 0;
    for (size_t i = 0; i < nsteps; ++i) { /* :56-104 */
        const uint64_t *s = steps + i * 66;
        if (s[13] == 1) continue;
        uint64_t v[6] = {s[6], s[7], s[8], s[9], s[12], s[25]};
        for (int k = 0; k < 6; ++k) T(8 + k, e) = gl_canon(v[k]);
        T(16, e) = 1;
        T(14, e) = prog_compress(v, beta);
        e++;
        if (s[26] == 1 || s[27] == MASK(OP_MLOAD) || s[27] == MASK(OP_MSTORE)) {
            uint64_t w[6] = {s[6], s[7], s[8], s[9], s[12] + 1, s[28]};
            for (int k = 0; k < 6; ++k) T(8 + k, e) = gl_canon(w[k]);
            T(16, e) = 1;
            T(14, e) = prog_compress(w, beta);
            e++;
        }
    }
    for (size_t j = 0; j < m; ++j) { /* :106-131 */
        const uint64_t *r = prog_rows + j * 6;
        for (int k = 0; k < 6; ++k) T(k, j) = gl_canon(r[k]);
        T(17, j) = 1;
        T(6, j) = prog_compress(r, beta);
    }
    orc_permuted_cols(&T(14, 0), &T(6, 0), n, &T(15, 0), &T(7, 0)); /* :132-135 */
#undef T
    if (beta_out) *beta_out = beta;
    return n;
}
