// Trace generation for the five small tables, completing the ola_generate_* family so that all
// twelve tables of generate_traces (circuits/src/generation/mod.rs:79-213) can be produced in HBM next to the prover.
//
//   generate_poseidon_chunk_trace   circuits/src/generation/poseidon_chunk.rs:7-88    PoseidonChunkRow records  -> [53][n]
//   generate_storage_access_trace   circuits/src/generation/storage.rs:7-123          StorageHashRow records    -> [48][n]
//   generate_tape_trace             circuits/src/generation/tape.rs:10-73             TapeRow records           -> [6][n]
//   generate_sccall_trace           circuits/src/generation/sccall.rs:11-64           SCCallRow records         -> [26][n]
//   generate_prog_chunk_trace       circuits/src/generation/prog.rs:158-249           program words             -> [40][n]
//
// The first four are one thread per table row (a record copied into its columns, the derived flag columns computed from the raw
// u64 values exactly as the Rust does, the padding rows of each table).  ProgChunk places the words of every program in lines
// of eight with a prefix sum over "this word starts a line", then ONE thread walks the lines in order, because each line's
// Poseidon input contains the previous line's output (a sponge in overwrite mode: prog.rs:206-224) -- the chain is sequential by
// construction, a few microseconds per line.
#include "common.h"
#include "gl.cuh"
#include "poseidon.cuh"

namespace ola {
namespace lookup {

// the reference's row count: the next power of two, at least 2 (every generator above opens with the same expression)
static inline bool rows_fit(size_t nrows, uint32_t log_n) { return log_n >= 1 && log_n <= 28 && nrows <= ((size_t)1 << log_n); }

// ---- PoseidonChunk ------------------------------------------------------------------------------------------------------------------
// record (32 u64, PoseidonChunkRow field order): 0 env_idx  1 clk  2 opcode  3 dst  4 op0  5 op1  6 acc_cnt  7..14 value[8]
// 15..18 cap[4]  19..30 hash[12]  31 is_ext_line.   columns: builtins/poseidon/columns.rs:42-68
__global__ void pchunk_fill_kernel(const uint64_t* __restrict__ rows, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t col[53];
#pragma unroll
    for (int c = 0; c < 53; ++c) col[c] = 0;
    if (i < nrows) {
        const uint64_t* r = rows + i * 32;
        const uint64_t op1 = r[5], acc = r[6], ext = r[31];
        col[1] = r[0], col[2] = (uint32_t)r[1], col[3] = r[2], col[4] = r[4], col[5] = op1, col[6] = r[3], col[7] = acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) col[8 + j] = r[7 + j];
#pragma unroll
        for (int j = 0; j < 4; ++j) col[16 + j] = r[15 + j];
#pragma unroll
        for (int j = 0; j < 12; ++j) col[20 + j] = r[19 + j];
        col[32] = ext;
        const bool result = op1 == acc;
        const int pad = result ? (int)(op1 % 8) : 0;
        col[33] = result ? 1 : 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) col[34 + j] = (pad == j) ? 1 : 0;
        col[42] = ext == 0 ? 1 : 0;
        if (ext == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) col[43 + j] = (pad != 0 && j >= pad) ? 0 : 1;
        }
        col[51] = ext;
    } else {
        col[52] = 1;
    }
#pragma unroll
    for (int c = 0; c < 53; ++c) out[(size_t)c * n + i] = c >= 33 && c != 51 ? col[c] : gl::canon(col[c]);
}
void poseidon_chunk_trace(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    OLA_CHECK(rows_fit(nrows, log_n), OLA_ERR_INVALID_ARG, "PoseidonChunk table: a power of two of at least 2 rows with room for every line");
    const size_t n = (size_t)1 << log_n;
    Launch lz(ctx, "gen_small_fill");
    pchunk_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_rows, nrows, n, d_out);
    check_launch("pchunk_fill_kernel");
}

// ---- StorageAccess ------------------------------------------------------------------------------------------------------------------
// record (38 u64, StorageHashRow field order): 0 storage_access_idx  1..4 pre_root  5..8 root  9 is_write  10 layer  11 layer_bit
// 12 addr_acc  13..16 addr  17..20 pre_path  21..24 path  25 hash_type  26..29 pre_hash  30..33 hash  34..37 sibling.
// The first n_access records are the storage accesses, the rest the program-hash reads (storage.rs:23).
__global__ void storage_fill_kernel(const uint64_t* __restrict__ rows, size_t n_access, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto put = [&](int c, uint64_t v) { out[(size_t)c * n + i] = gl::canon(v); };
    if (i < nrows) {
        const uint64_t* r = rows + i * 38;
        const uint64_t layer = r[10], bit = r[11];
        put(0, r[0]);
        for (int j = 0; j < 4; ++j) {
            put(1 + j, r[1 + j]), put(5 + j, r[5 + j]), put(13 + j, r[13 + j]), put(17 + j, r[17 + j]), put(21 + j, r[21 + j]);
            put(25 + j, r[34 + j]), put(30 + j, r[26 + j]), put(34 + j, r[30 + j]);
        }
        put(9, r[9]), put(10, layer), put(11, bit), put(12, r[12]), put(29, r[25]);
        put(38, layer == 1), put(39, layer == 64), put(40, layer == 128), put(41, layer == 192), put(42, layer == 256);
        put(43, layer < 64 ? 1 : layer < 128 ? 2 : layer < 192 ? 3 : layer < 256 ? 4 : layer == 256 ? 5 : 0);
        put(44, bit == 0), put(45, bit == 1);
        put(46, i >= n_access && layer == 256);
        put(47, 0);
    } else {
        for (int c = 0; c < 48; ++c) put(c, 0);
        if (nrows)
            for (int j = 0; j < 4; ++j) put(5 + j, rows[(nrows - 1) * 38 + 5 + j]);  // the last root carried on
        put(47, 1);
    }
}
void storage_access_trace(ola_ctx* ctx, const uint64_t* d_rows, size_t n_access, size_t n_prog, uint32_t log_n, uint64_t* d_out) {
    OLA_CHECK(rows_fit(n_access + n_prog, log_n), OLA_ERR_INVALID_ARG, "StorageAccess table: a power of two of at least 2 rows with room for every layer row");
    const size_t n = (size_t)1 << log_n;
    Launch lz(ctx, "gen_small_fill");
    storage_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_rows, n_access, n_access + n_prog, n, d_out);
    check_launch("storage_fill_kernel");
}

// ---- Tape -----------------------------------------------------------------------------------------------------------------------------
// record (5 u64, TapeRow field order): is_init  opcode  addr  value  filter_looked.  Padding repeats the last row as an unlooked
// TLOAD (tape.rs:33-65); OlaOpcode::TLOAD.binary_bit_mask() = 1 << 9.
__global__ void tape_fill_kernel(const uint64_t* __restrict__ rows, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v[6] = {0, 0, 0, 0, 0, 0};
    if (i < nrows) {
        const uint64_t* r = rows + i * 5;
        v[1] = r[0] ? 1 : 0, v[2] = gl::canon(r[1]), v[3] = gl::canon(r[2]), v[4] = gl::canon(r[3]), v[5] = gl::canon(r[4]);
    } else {
        v[2] = 1ull << 9;
        if (nrows) {
            const uint64_t* r = rows + (nrows - 1) * 5;
            v[1] = r[0] ? 1 : 0, v[3] = gl::canon(r[2]), v[4] = gl::canon(r[3]);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) out[(size_t)c * n + i] = v[c];
}
void tape_trace(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    OLA_CHECK(rows_fit(nrows, log_n), OLA_ERR_INVALID_ARG, "Tape table: a power of two of at least 2 rows with room for every cell");
    const size_t n = (size_t)1 << log_n;
    Launch lz(ctx, "gen_small_fill");
    tape_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_rows, nrows, n, d_out);
    check_launch("tape_fill_kernel");
}

// ---- SCCall -------------------------------------------------------------------------------------------------------------------------
// record (24 u64, SCCallRow field order): 0 caller_env_idx  1..4 addr_storage  5..8 addr_code  9 caller_op1_imm  10 clk_caller_call
// 11 clk_caller_ret  12..21 regs[10]  22 callee_env_idx  23 clk_callee_end.  Column c + 1 = record word c (sccall/columns.rs:4-20),
// column 0 (tx_idx) is 0 and column 25 marks the padding rows.
__global__ void sccall_fill_kernel(const uint64_t* __restrict__ rows, size_t nrows, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = 0;
#pragma unroll
    for (int c = 0; c < 24; ++c) out[(size_t)(c + 1) * n + i] = i < nrows ? gl::canon(rows[i * 24 + c]) : 0;
    out[(size_t)25 * n + i] = i < nrows ? 0 : 1;
}
void sccall_trace(ola_ctx* ctx, const uint64_t* d_rows, size_t nrows, uint32_t log_n, uint64_t* d_out) {
    OLA_CHECK(rows_fit(nrows, log_n), OLA_ERR_INVALID_ARG, "SCCall table: a power of two of at least 2 rows with room for every call");
    const size_t n = (size_t)1 << log_n;
    Launch lz(ctx, "gen_small_fill");
    sccall_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_rows, nrows, n, d_out);
    check_launch("sccall_fill_kernel");
}

// ---- ProgChunk ------------------------------------------------------------------------------------------------------------------------
// prog_rows [m][6] = (code address 0..3, pc, word) for every word of every program, in the order the Rust walks `progs` (the same
// buffer ola_generate_program_trace takes); a program starts where pc == 0, a line where pc % 8 == 0.
__global__ void pc_flag_kernel(const uint64_t* __restrict__ rows, size_t m, uint32_t* __restrict__ starts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) starts[i] = (rows[i * 6 + 4] % 8 == 0) ? 1u : 0u;
}
// line index of word i = (number of line starts among words 0..i) - 1 = at[i] + starts[i] - 1
__global__ void pc_place_kernel(const uint64_t* __restrict__ rows, size_t m, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ at,
                                size_t n, uint64_t* __restrict__ out, uint32_t* __restrict__ line_len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t* r = rows + i * 6;
    if (at[i] + starts[i] == 0) return;  // words in front of the first line start (a program that does not begin at pc = 0): no line to put them in
    const size_t line = (size_t)at[i] + starts[i] - 1;
    const uint64_t pc = r[4];
    const int j = (int)(pc % 8);
    out[(size_t)(5 + j) * n + line] = gl::canon(r[5]);
    out[(size_t)(31 + j) * n + line] = 1;
    if (j == 0) {
        for (int k = 0; k < 4; ++k) out[(size_t)k * n + line] = gl::canon(r[k]);
        out[(size_t)4 * n + line] = pc;
        out[(size_t)29 * n + line] = pc == 0 ? 1 : 0;
    }
    const bool last_of_program = i + 1 == m || rows[(i + 1) * 6 + 4] == 0;
    const bool last_of_line = last_of_program || j == 7;
    if (last_of_line) line_len[line] = (uint32_t)(j + 1);
    if (last_of_program) out[(size_t)30 * n + line] = 1;
}
// one thread: the sponge over the lines, state carried from line to line -- and, as in the reference, from program to program
// (prog.rs:199 declares pre_hash once, outside the loop over all lines)
__global__ void pc_chain_kernel(size_t nlines, const uint32_t* __restrict__ line_len, size_t n, uint64_t* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t pre[12];
    for (int j = 0; j < 12; ++j) pre[j] = 0;
    for (size_t line = 0; line < nlines; ++line) {
        const int len = (int)line_len[line];
        uint64_t s[12];
        for (int j = 0; j < 8; ++j) {
            if (j < len) {
                s[j] = out[(size_t)(5 + j) * n + line];
            } else {
                s[j] = pre[j];
                out[(size_t)(5 + j) * n + line] = pre[j];
            }
        }
        for (int j = 0; j < 4; ++j) {
            s[8 + j] = pre[8 + j];
            out[(size_t)(13 + j) * n + line] = pre[8 + j];
        }
        poseidon::permute<3>(s);
        for (int j = 0; j < 12; ++j) {
            pre[j] = s[j];
            out[(size_t)(17 + j) * n + line] = s[j];
        }
    }
}
__global__ void pc_padding_kernel(size_t nlines, size_t n, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)39 * n + i] = i < nlines ? 0 : 1;
}
void prog_chunk_trace(ola_ctx* ctx, const uint64_t* d_prog_rows, size_t m, uint32_t log_n, uint64_t* d_out) {
    OLA_CHECK(log_n >= 1 && log_n <= 28 && m <= ((size_t)1 << 31), OLA_ERR_INVALID_ARG, "ProgChunk table: a power of two of at least 2 rows");
    const size_t n = (size_t)1 << log_n;
    OLA_CUDA(cudaMemsetAsync(d_out, 0, 40 * n * sizeof(uint64_t), ctx->stream));
    uint32_t nlines = 0;
    if (m) {
        const size_t nb = (m + SCAN_BLOCK - 1) / SCAN_BLOCK;
        Buf w32((2 * m + nb + 8) / 2 + 4), len32(n / 2 + 1);
        uint32_t* starts = w32.u32();
        uint32_t* at = starts + m;
        uint32_t* sums = at + m;
        uint32_t* total = sums + nb;
        Launch lz(ctx, "gen_small_fill");
        const unsigned gs = (unsigned)((m + 255) / 256);
        pc_flag_kernel<<<gs, 256, 0, ctx->stream>>>(d_prog_rows, m, starts);
        exclusive_scan(ctx, starts, at, sums, total, m);
        OLA_CUDA(cudaMemcpyAsync(&nlines, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
        OLA_CHECK(nlines <= n, OLA_ERR_INVALID_ARG, "ProgChunk table: the programs have more lines of eight words than the table has rows");
        pc_place_kernel<<<gs, 256, 0, ctx->stream>>>(d_prog_rows, m, starts, at, n, d_out, len32.u32());
        pc_chain_kernel<<<1, 32, 0, ctx->stream>>>(nlines, len32.u32(), n, d_out);
        check_launch("pc_chain_kernel");
        OLA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    Launch lz(ctx, "gen_small_fill");
    pc_padding_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(nlines, n, d_out);
    check_launch("pc_padding_kernel");
}

}  // namespace lookup
}  // namespace ola
