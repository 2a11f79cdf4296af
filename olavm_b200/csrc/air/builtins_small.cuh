// AIR constraints of the small builtin tables, one device function per table, constraints emitted in the
// reference's source order (the consumer is Horner in alpha).
//   Cmp         circuits/src/builtins/cmp/cmp_stark.rs:21-45        (columns.rs:16-22)
//   RangeCheck  circuits/src/builtins/rangecheck/rangecheck_stark.rs:27-67 (columns.rs:25-39)
#pragma once
#include "air_common.cuh"

namespace ola {
namespace air {

// CpuStark: circuits/src/cpu/cpu_stark.rs:871-946 (shared transcription in cpu_air.h)
struct Cpu {
    enum { COLUMNS = cpu::NUM_CPU_COLS };
    static constexpr int CONSTRAINT_DEGREE = 7;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc) { cpu::eval<Fp, Row, Consumer>(lv, nv, yc); }
};

// MemoryStark: circuits/src/memory/memory_stark.rs:92-340 (shared transcription in mem_air.h)
struct Memory {
    enum { COLUMNS = mem::NUM_MEM_COLS };
    static constexpr int CONSTRAINT_DEGREE = 8;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc) { mem::eval<Fp, Row, Consumer>(lv, nv, yc); }
};

struct Cmp {
    enum { OP0 = 0, OP1, GTE, ABS_DIFF, ABS_DIFF_INV, FILTER_LOOKING_RC, COLUMNS };
    static constexpr int CONSTRAINT_DEGREE = 3;
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc) {
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
    static __device__ __forceinline__ void eval(const Row& lv, const Row& nv, Consumer& yc) {
        Fp val = lv[VAL], limb_lo = lv[LIMB_LO], limb_hi = lv[LIMB_HI];
        Fp sum = limb_lo + limb_hi * fp(1 << 16);
        yc.constraint(val - sum);
        eval_lookups(lv, nv, yc, LIMB_LO_PERMUTED, FIX_RANGE_CHECK_U16_PERMUTED_LO);
        eval_lookups(lv, nv, yc, LIMB_HI_PERMUTED, FIX_RANGE_CHECK_U16_PERMUTED_HI);
    }
};

}  // namespace air
}  // namespace ola
