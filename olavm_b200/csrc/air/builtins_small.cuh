// AIR constraints of the small builtin tables, one device function per table, constraints emitted in the
// reference's source order (the consumer is Horner in alpha).
//   Cmp         circuits/src/builtins/cmp/cmp_stark.rs:21-45        (columns.rs:16-22)
//   RangeCheck  circuits/src/builtins/rangecheck/rangecheck_stark.rs:27-67 (columns.rs:25-39)
#pragma once
#include "../poseidon.cuh"
#include "air_common.cuh"
#include "hash_air.h"

namespace ola {
namespace air {

// CpuStark: circuits/src/cpu/cpu_stark.rs:871-946 (shared transcription in cpu_air.h)
struct Cpu {
    enum { COLUMNS = cpu::NUM_CPU_COLS };
    static constexpr int CONSTRAINT_DEGREE = 7;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp) { cpu::eval<Fp, Row, Consumer>(lv, nv, yc); }
};

// MemoryStark: circuits/src/memory/memory_stark.rs:92-340 (shared transcription in mem_air.h)
struct Memory {
    enum { COLUMNS = mem::NUM_MEM_COLS };
    static constexpr int CONSTRAINT_DEGREE = 8;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp) { mem::eval<Fp, Row, Consumer>(lv, nv, yc); }
};

struct Cmp {
    enum { OP0 = 0, OP1, GTE, ABS_DIFF, ABS_DIFF_INV, FILTER_LOOKING_RC, COLUMNS };
    static constexpr int CONSTRAINT_DEGREE = 3;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp) {
        Fp op0 = lv[OP0], op1 = lv[OP1], gte = lv[GTE], abs_diff = lv[ABS_DIFF], abs_diff_inv = lv[ABS_DIFF_INV];
        // gte must be binary
        yc.constraint(gte * (one() - gte));
        // abs_diff calculation
        yc.constraint(gte * (op0 - op1 - abs_diff));
        yc.constraint((one() - gte) * (op1 - op0 - abs_diff));
        // abs_diff * abs_diff_inv = 1 when gte = 0
        yc.constraint((one() - gte) * (one() - abs_diff * abs_diff_inv));
    }
};

struct RangeCheck {
    enum { CPU_FILTER = 0, MEMORY_SORT_FILTER, MEMORY_REGION_FILTER, CMP_FILTER, VAL, LIMB_LO, LIMB_HI, LIMB_LO_PERMUTED, LIMB_HI_PERMUTED,
           FIX_RANGE_CHECK_U16, FIX_RANGE_CHECK_U16_PERMUTED_LO, FIX_RANGE_CHECK_U16_PERMUTED_HI, COLUMNS };
    static constexpr int CONSTRAINT_DEGREE = 3;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp) {
        Fp val = lv[VAL], limb_lo = lv[LIMB_LO], limb_hi = lv[LIMB_HI];
        Fp sum = limb_lo + limb_hi * fp(1 << 16);
        yc.constraint(val - sum);
        eval_lookups(lv, nv, yc, LIMB_LO_PERMUTED, FIX_RANGE_CHECK_U16_PERMUTED_LO);
        eval_lookups(lv, nv, yc, LIMB_HI_PERMUTED, FIX_RANGE_CHECK_U16_PERMUTED_HI);
    }
};

// ---- tables whose constraint bodies are the shared transcriptions of builtins_air.h / hash_air.h ----
#define OLA_AIR_WRAPPER(NAME, NS, NCOLS, DEGREE)                                                                   \
    struct NAME {                                                                                                  \
        enum { COLUMNS = NS::NCOLS };                                                                              \
        static constexpr int CONSTRAINT_DEGREE = DEGREE;                                                           \
        static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp) { NS::eval<Fp, Row, Consumer>(lv, nv, yc); } \
    };
OLA_AIR_WRAPPER(Tape, tape, NUM_COL_TAPE, 5)                           // tape_stark.rs:139-141
OLA_AIR_WRAPPER(SCCall, sccall, NUM_COL_SCCALL, 1)                     // sccall_stark.rs:96-98
OLA_AIR_WRAPPER(ProgChunk, prog_chunk, NUM_PROG_CHUNK_COLS, 4)         // prog_chunk_stark.rs:172-174
OLA_AIR_WRAPPER(StorageAccess, storage, NUM_COL_ST, 4)                 // storage_access_stark.rs:330-332
OLA_AIR_WRAPPER(PoseidonChunk, psdn_chunk, NUM_POSEIDON_CHUNK_COLS, 3) // poseidon_chunk_stark.rs:287-289
#undef OLA_AIR_WRAPPER

// compress-challenge tables: beta comes from trace generation (generation/mod.rs:183-188) through ola_prove
struct Program {
    enum { COLUMNS = program::NUM_PROG_COLS };
    static constexpr int CONSTRAINT_DEGREE = 3;  // program_stark.rs:107-109
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp beta) { program::eval<Fp, Row, Consumer>(lv, nv, yc, beta); }
};
struct Bitwise {
    enum { COLUMNS = bitwise::COL_NUM_BITWISE };
    static constexpr int CONSTRAINT_DEGREE = 3;  // bitwise_stark.rs:346-348
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp beta) { bitwise::eval<Fp, Row, Consumer>(lv, nv, yc, beta); }
};

// Poseidon parameter tables for the Poseidon table's AIR: the hash kernels' __constant__ copies
struct PoseidonParams {
    static __device__ __forceinline__ uint64_t round(int i) { return poseidon::c_round[i]; }
    static __device__ __forceinline__ uint64_t circ(int i) {
        constexpr uint64_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
        return C[i];
    }
    static __device__ __forceinline__ uint64_t diag(int i) { return i == 0 ? 8 : 0; }
    static __device__ __forceinline__ uint64_t first(int i) { return poseidon::c_first[i]; }
    static __device__ __forceinline__ uint64_t partial(int r) { return poseidon::c_partial[r]; }
    static __device__ __forceinline__ uint64_t init(int r, int c) { return poseidon::c_init[r * 11 + c]; }
    static __device__ __forceinline__ uint64_t what(int r, int i) { return poseidon::c_whats[r * 11 + i]; }
    static __device__ __forceinline__ uint64_t vs(int r, int i) { return poseidon::c_vs[r * 11 + i]; }
};
struct Poseidon {
    enum { COLUMNS = psdn::NUM_POSEIDON_COLS };
    static constexpr int CONSTRAINT_DEGREE = 7;  // poseidon_stark.rs:145-147
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc, Fp) { psdn::eval<Fp, Row, Consumer, PoseidonParams>(lv, nv, yc); }
};

}  // namespace air
}  // namespace ola
