/* ORACLE (test infrastructure, NOT product code): the generators of the five small tables restated, statement for statement,
 * as the serial loops the reference runs.
 *
 *   orc_generate_poseidon_chunk_trace   circuits/src/generation/poseidon_chunk.rs:7-88   columns: builtins/poseidon/columns.rs:42-68
 *   orc_generate_storage_access_trace   circuits/src/generation/storage.rs:7-123         columns: builtins/storage/columns.rs:3-33
 *   orc_generate_tape_trace             circuits/src/generation/tape.rs:10-73            columns: builtins/tape/columns.rs:3-9
 *   orc_generate_sccall_trace           circuits/src/generation/sccall.rs:11-64          columns: builtins/sccall/columns.rs:4-20
 *   orc_generate_prog_chunk_trace       circuits/src/generation/prog.rs:158-249          columns: program/columns.rs:47-62
 *
 * Records are the Rust structs of core/src/trace/trace.rs flattened in field order (layouts at each function); tables are
 * column-major out[ncols][n].  Every function returns the reference's row count (next power of two, at least 2) and fills `out`
 * only when out_rows >= that count (out_rows is then the table's n: the reference's count, or a larger power of two the caller
 * pads to, which the reference would reach with more padding rows). */
#include <string.h>

#include "oracle.h"

static size_t padded_rows(size_t filled) { /* "if !len.is_power_of_two() || len < 2 { if len < 2 { 2 } else { next_power_of_two } }" */
    size_t n = 2;
    while (n < filled) n <<= 1;
    return n;
}
#define T(c, i) out[(size_t)(c) * n + (i)]

/* cells [k][32]: 0 env_idx  1 clk  2 opcode  3 dst  4 op0  5 op1  6 acc_cnt  7..14 value[8]  15..18 cap[4]  19..30 hash[12]  31 is_ext_line */
size_t orc_generate_poseidon_chunk_trace(const uint64_t *cells, size_t k, uint64_t *out, size_t out_rows) {
    const size_t need = padded_rows(k);
    if (!out || out_rows < need) return need;
    const size_t n = out_rows;
    memset(out, 0, 53 * n * sizeof(uint64_t));
    enum { TX = 0, ENV, CLK, OPCODE, OP0, OP1, DST, ACC, VALUE = 8, CAP = 16, HASH = 20, IS_EXT = 32, IS_RESULT = 33, FIRST_PAD = 34, LOOKED_CPU = 42,
           LOOKING_MEM = 43, LOOKING_PSDN = 51, IS_PAD = 52 };
    for (size_t i = 0; i < k; ++i) { /* :21-74 */
        const uint64_t *c = cells + i * 32;
        T(TX, i) = 0;
        T(ENV, i) = gl_canon(c[0]);
        T(CLK, i) = (uint32_t)c[1];
        T(OPCODE, i) = gl_canon(c[2]);
        T(OP0, i) = gl_canon(c[4]);
        T(OP1, i) = gl_canon(c[5]);
        T(DST, i) = gl_canon(c[3]);
        T(ACC, i) = gl_canon(c[6]);
        for (int j = 0; j < 8; ++j) T(VALUE + j, i) = gl_canon(c[7 + j]);
        for (int j = 0; j < 4; ++j) T(CAP + j, i) = gl_canon(c[15 + j]);
        for (int j = 0; j < 12; ++j) T(HASH + j, i) = gl_canon(c[19 + j]);
        T(IS_EXT, i) = gl_canon(c[31]);
        T(IS_RESULT, i) = c[5] == c[6] ? 1 : 0;
        if (c[5] == c[6]) {
            const int first_padding_index = (int)(c[5] % 8);
            if (first_padding_index != 0) T(FIRST_PAD + first_padding_index, i) = 1;
        }
        T(LOOKED_CPU, i) = c[31] == 0 ? 1 : 0;
        if (c[31] == 1) {
            for (int j = 0; j < 8; ++j) T(LOOKING_MEM + j, i) = 1;
            if (c[5] == c[6]) {
                const int first_padding_index = (int)(c[5] % 8);
                if (first_padding_index != 0)
                    for (int j = first_padding_index; j < 8; ++j) T(LOOKING_MEM + j, i) = 0;
            }
        }
        T(LOOKING_PSDN, i) = gl_canon(c[31]);
    }
    for (size_t i = k; i < n; ++i) T(IS_PAD, i) = 1; /* :76-80 */
    return need;
}

/* rows [n_access + n_prog][38]: 0 storage_access_idx  1..4 pre_root  5..8 root  9 is_write  10 layer  11 layer_bit  12 addr_acc  13..16 addr
 * 17..20 pre_path  21..24 path  25 hash_type  26..29 pre_hash  30..33 hash  34..37 sibling; the accesses first, then the program-hash reads */
size_t orc_generate_storage_access_trace(const uint64_t *rows, size_t n_access, size_t n_prog, uint64_t *out, size_t out_rows) {
    const size_t k = n_access + n_prog, need = padded_rows(k);
    if (!out || out_rows < need) return need;
    const size_t n = out_rows;
    memset(out, 0, 48 * n * sizeof(uint64_t));
    enum { IDX = 0, PRE_ROOT = 1, ROOT = 5, IS_WRITE = 9, LAYER, LAYER_BIT, ADDR_ACC, ADDR = 13, PRE_PATH = 17, PATH = 21, SIB = 25, HASH_TYPE = 29,
           PRE_HASH = 30, HASH = 34, IS_L1 = 38, IS_L64, IS_L128, IS_L192, IS_L256, MARKER, BIT0, BIT1, FOR_PROG, IS_PAD };
    for (size_t i = 0; i < k; ++i) { /* :23-83 */
        const uint64_t *c = rows + i * 38;
        const uint64_t layer = c[10], layer_bit = c[11];
        T(IDX, i) = gl_canon(c[0]);
        for (int j = 0; j < 4; ++j) T(PRE_ROOT + j, i) = gl_canon(c[1 + j]);
        for (int j = 0; j < 4; ++j) T(ROOT + j, i) = gl_canon(c[5 + j]);
        T(IS_WRITE, i) = gl_canon(c[9]);
        T(LAYER, i) = gl_canon(layer);
        T(LAYER_BIT, i) = gl_canon(layer_bit);
        T(ADDR_ACC, i) = gl_canon(c[12]);
        for (int j = 0; j < 4; ++j) T(ADDR + j, i) = gl_canon(c[13 + j]);
        for (int j = 0; j < 4; ++j) T(PRE_PATH + j, i) = gl_canon(c[17 + j]);
        for (int j = 0; j < 4; ++j) T(PATH + j, i) = gl_canon(c[21 + j]);
        for (int j = 0; j < 4; ++j) T(SIB + j, i) = gl_canon(c[34 + j]);
        T(HASH_TYPE, i) = gl_canon(c[25]);
        for (int j = 0; j < 4; ++j) T(PRE_HASH + j, i) = gl_canon(c[26 + j]);
        for (int j = 0; j < 4; ++j) T(HASH + j, i) = gl_canon(c[30 + j]);
        T(IS_L1, i) = layer == 1;
        T(IS_L64, i) = layer == 64;
        T(IS_L128, i) = layer == 128;
        T(IS_L192, i) = layer == 192;
        T(IS_L256, i) = layer == 256;
        if (layer < 64)
            T(MARKER, i) = 1;
        else if (layer < 128)
            T(MARKER, i) = 2;
        else if (layer < 192)
            T(MARKER, i) = 3;
        else if (layer < 256)
            T(MARKER, i) = 4;
        else if (layer == 256)
            T(MARKER, i) = 5;
        else
            T(MARKER, i) = 0;
        T(BIT0, i) = layer_bit == 0;
        T(BIT1, i) = layer_bit == 1;
        if (i < n_access)
            T(FOR_PROG, i) = 0;
        else if (layer == 256)
            T(FOR_PROG, i) = 1;
        else
            T(FOR_PROG, i) = 0;
        T(IS_PAD, i) = 0;
    }
    uint64_t last_root[4] = {0, 0, 0, 0}; /* :85-107 */
    if (k != 0)
        for (int j = 0; j < 4; ++j) last_root[j] = T(ROOT + j, k - 1);
    for (size_t i = k; i < n; ++i) { /* :108-115 */
        for (int j = 0; j < 4; ++j) T(ROOT + j, i) = last_root[j];
        T(IS_PAD, i) = 1;
    }
    return need;
}

/* cells [k][5]: is_init  opcode  addr  value  filter_looked */
size_t orc_generate_tape_trace(const uint64_t *cells, size_t k, uint64_t *out, size_t out_rows) {
    const size_t need = padded_rows(k);
    if (!out || out_rows < need) return need;
    const size_t n = out_rows;
    memset(out, 0, 6 * n * sizeof(uint64_t));
    enum { TX = 0, IS_INIT, OPCODE, ADDR, VALUE, LOOKED };
    for (size_t i = 0; i < k; ++i) { /* :23-30 */
        const uint64_t *c = cells + i * 5;
        T(TX, i) = 0;
        T(IS_INIT, i) = c[0] ? 1 : 0;
        T(OPCODE, i) = gl_canon(c[1]);
        T(ADDR, i) = gl_canon(c[2]);
        T(VALUE, i) = gl_canon(c[3]);
        T(LOOKED, i) = gl_canon(c[4]);
    }
    const uint64_t last_tx_idx = k == 0 ? 0 : T(TX, k - 1); /* :32-51 */
    const uint64_t last_is_init = k == 0 ? 0 : T(IS_INIT, k - 1);
    const uint64_t last_addr = k == 0 ? 0 : T(ADDR, k - 1);
    const uint64_t last_value = k == 0 ? 0 : T(VALUE, k - 1);
    const uint64_t op_tload = 1ull << 9; /* OlaOpcode::TLOAD.binary_bit_mask(), core/src/vm/opcodes.rs */
    for (size_t i = k; i < n; ++i) {     /* :55-64 */
        T(TX, i) = last_tx_idx;
        T(IS_INIT, i) = last_is_init;
        T(OPCODE, i) = op_tload;
        T(ADDR, i) = last_addr;
        T(VALUE, i) = last_value;
        T(LOOKED, i) = 0;
    }
    return need;
}

/* cells [k][24]: 0 caller_env_idx  1..4 addr_storage  5..8 addr_code  9 caller_op1_imm  10 clk_caller_call  11 clk_caller_ret  12..21 regs[10]
 * 22 callee_env_idx  23 clk_callee_end */
size_t orc_generate_sccall_trace(const uint64_t *cells, size_t k, uint64_t *out, size_t out_rows) {
    const size_t need = padded_rows(k);
    if (!out || out_rows < need) return need;
    const size_t n = out_rows;
    memset(out, 0, 26 * n * sizeof(uint64_t));
    enum { TX = 0, CALLER_ENV = 1, EXE_CTX = 2, CODE_CTX = 6, OP1_IMM = 10, CLK_CALL, CLK_RET, REGS = 13, CALLEE_ENV = 23, CLK_END, IS_PAD };
    for (size_t i = 0; i < k; ++i) { /* :24-49 */
        const uint64_t *c = cells + i * 24;
        T(TX, i) = 0;
        T(CALLER_ENV, i) = gl_canon(c[0]);
        for (int j = 0; j < 4; ++j) T(EXE_CTX + j, i) = gl_canon(c[1 + j]);
        for (int j = 0; j < 4; ++j) T(CODE_CTX + j, i) = gl_canon(c[5 + j]);
        T(OP1_IMM, i) = gl_canon(c[9]);
        T(CLK_CALL, i) = gl_canon(c[10]);
        T(CLK_RET, i) = gl_canon(c[11]);
        for (int j = 0; j < 10; ++j) T(REGS + j, i) = gl_canon(c[12 + j]);
        T(CALLEE_ENV, i) = gl_canon(c[22]);
        T(CLK_END, i) = gl_canon(c[23]);
    }
    for (size_t i = k; i < n; ++i) T(IS_PAD, i) = 1; /* :50-54 */
    return need;
}

/* prog_rows [m][6] = (code address 0..3, pc, word) for every word of every program, programs in the order of `progs`; a program
 * begins where pc == 0 */
size_t orc_generate_prog_chunk_trace(const uint64_t *prog_rows, size_t m, uint64_t *out, size_t out_rows) {
    /* :161-184: (addr, chunk_idx * 8, chunk, is_first_line, is_result_line) for insts.chunks(8) of every program */
    size_t lines = 0;
    for (size_t i = 0; i < m; ++i) lines += prog_rows[i * 6 + 4] % 8 == 0;
    const size_t need = padded_rows(lines);
    if (!out || out_rows < need) return need;
    const size_t n = out_rows;
    memset(out, 0, 40 * n * sizeof(uint64_t));
    enum { ADDR = 0, START_PC = 4, INST = 5, CAP = 13, HASH = 17, IS_FIRST = 29, IS_RESULT = 30, LOOKING_PROG = 31, IS_PAD = 39 };
    uint64_t pre_hash[12] = {0}; /* :199 */
    size_t i = 0, w = 0;
    while (w < m) { /* :200-237 */
        const uint64_t *r = prog_rows + w * 6;
        const uint64_t start_pc = r[4];
        size_t chunk_len = 1; /* the words of this line: up to eight, ending with the program */
        while (chunk_len < 8 && w + chunk_len < m && prog_rows[(w + chunk_len) * 6 + 4] != 0) ++chunk_len;
        const int is_first_line = start_pc == 0;
        const int is_result_line = w + chunk_len == m || prog_rows[(w + chunk_len) * 6 + 4] == 0;
        for (int j = 0; j < 4; ++j) T(ADDR + j, i) = gl_canon(r[j]);
        T(START_PC, i) = start_pc;
        uint64_t hash_input[12];
        for (size_t j = 0; j < chunk_len; ++j) {
            const uint64_t v = gl_canon(prog_rows[(w + j) * 6 + 5]);
            T(INST + j, i) = v;
            hash_input[j] = v;
        }
        for (size_t j = chunk_len; j < 8; ++j) {
            T(INST + j, i) = pre_hash[j];
            hash_input[j] = pre_hash[j];
        }
        for (int j = 0; j < 4; ++j) {
            T(CAP + j, i) = pre_hash[j + 8];
            hash_input[j + 8] = pre_hash[j + 8];
        }
        orc_poseidon(hash_input); /* calculate_poseidon */
        for (int j = 0; j < 12; ++j) {
            T(HASH + j, i) = hash_input[j];
            pre_hash[j] = hash_input[j];
        }
        T(IS_FIRST, i) = is_first_line;
        T(IS_RESULT, i) = is_result_line;
        for (size_t j = 0; j < chunk_len; ++j) T(LOOKING_PROG + j, i) = 1;
        w += chunk_len;
        ++i;
    }
    for (size_t r = lines; r < n; ++r) T(IS_PAD, r) = 1; /* :239-243 */
    return need;
}
