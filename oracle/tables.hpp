/* ORACLE (test infrastructure, NOT product code) -- CPU restatement of the AIR tables ("chips") and of the
 * cross-table-lookup registry.  Each eval<> follows the Rust eval_packed_generic of the cited file line by
 * line and emits the constraints in source order (the consumer is Horner in alpha, so order is semantics:
 * circuits/src/stark/constraint_consumer.rs:60-65).
 *
 * Tables restated so far (enum order of circuits/src/stark/ola_stark.rs:104-119):
 *   3 Cmp         circuits/src/builtins/cmp/{columns.rs:16-22, cmp_stark.rs:21-45, :88-108}
 *   4 RangeCheck  circuits/src/builtins/rangecheck/{columns.rs:25-39, rangecheck_stark.rs:27-108, :111-140},
 *                 circuits/src/stark/lookup.rs:13-35
 */
#ifndef ORC_TABLES_HPP
#define ORC_TABLES_HPP
#include "stark.hpp"
/* The CPU table's constraint body and the CTL registry are a single transcription shared with the product
 * (olavm_b200/csrc/air/{cpu_air,ctl_registry}.h); see the note at the top of cpu_air.h and DESIGN.md section 4. */
#include "../olavm_b200/csrc/air/ctl_registry.h"

namespace orc {

enum TableId { T_CPU = 0, T_MEMORY, T_BITWISE, T_CMP, T_RANGECHECK, T_POSEIDON, T_POSEIDON_CHUNK, T_STORAGE, T_TAPE, T_SCCALL, T_PROGRAM, T_PROG_CHUNK, T_NUM };

/* lookup.rs:13-35 */
template <class O>
void eval_lookups(const P<O>* lv, const P<O>* nv, Consumer<O>& yc, int col_permuted_input, int col_permuted_table) {
    P<O> local_perm_input = lv[col_permuted_input];
    P<O> next_perm_table = nv[col_permuted_table];
    P<O> next_perm_input = nv[col_permuted_input];
    P<O> diff_input_prev = next_perm_input - local_perm_input;
    P<O> diff_input_table = next_perm_input - next_perm_table;
    yc.constraint(diff_input_prev * diff_input_table);
    yc.constraint_last_row(diff_input_table);
}

/* ---- Cmp (cmp_stark.rs:21-45) ---- */
namespace cmp {
enum { OP0 = 0, OP1, GTE, ABS_DIFF, ABS_DIFF_INV, FILTER_LOOKING_RC, NUM };
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    T op0 = lv[OP0], op1 = lv[OP1], gte = lv[GTE], abs_diff = lv[ABS_DIFF], abs_diff_inv = lv[ABS_DIFF_INV];
    yc.constraint(gte * (T::one() - gte));
    yc.constraint(gte * (op0 - op1 - abs_diff));
    yc.constraint((T::one() - gte) * (op1 - op0 - abs_diff));
    yc.constraint((T::one() - gte) * (T::one() - abs_diff * abs_diff_inv));
}
}  // namespace cmp

/* ---- RangeCheck (rangecheck_stark.rs:27-108) ---- */
namespace rangecheck {
enum { CPU_FILTER = 0, MEMORY_SORT_FILTER, MEMORY_REGION_FILTER, CMP_FILTER, VAL, LIMB_LO, LIMB_HI, LIMB_LO_PERMUTED, LIMB_HI_PERMUTED,
       FIX_RANGE_CHECK_U16, FIX_RANGE_CHECK_U16_PERMUTED_LO, FIX_RANGE_CHECK_U16_PERMUTED_HI, NUM };
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    T val = lv[VAL], limb_lo = lv[LIMB_LO], limb_hi = lv[LIMB_HI];
    T sum = limb_lo + limb_hi * T::c(1 << 16);
    yc.constraint(val - sum);
    eval_lookups<O>(lv, nv, yc, LIMB_LO_PERMUTED, FIX_RANGE_CHECK_U16_PERMUTED_LO);
    eval_lookups<O>(lv, nv, yc, LIMB_HI_PERMUTED, FIX_RANGE_CHECK_U16_PERMUTED_HI);
}
}  // namespace rangecheck

}  // namespace orc
namespace ola {
namespace air {
template <> inline orc::P<orc::FOps> kc<orc::P<orc::FOps>>(uint64_t k) { return orc::P<orc::FOps>::c(k); }
template <> inline orc::P<orc::EOps> kc<orc::P<orc::EOps>>(uint64_t k) { return orc::P<orc::EOps>::c(k); }
template <> inline bool is_zero<orc::P<orc::FOps>>(const orc::P<orc::FOps>& x) { return x.v == 0; }
template <> inline bool is_zero<orc::P<orc::EOps>>(const orc::P<orc::EOps>& x) { return x.v.c0 == 0 && x.v.c1 == 0; }
}  // namespace air
}  // namespace ola
namespace orc {
/* ---- Cpu (cpu_stark.rs:871-946, shared transcription) ---- */
namespace cpu_t {
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    ola::air::cpu::eval<P<O>, const P<O>*, Consumer<O>>(lv, nv, yc);
}
}  // namespace cpu_t
/* ---- Memory (memory_stark.rs:92-340, shared transcription) ---- */
namespace mem_t {
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    ola::air::mem::eval<P<O>, const P<O>*, Consumer<O>>(lv, nv, yc);
}
}  // namespace mem_t

template <class EvalB, class EvalE>
Table make_table(const char* name, int cols, int degree, EvalB eb, EvalE ee, std::vector<PermutationPair> pp = {}) {
    Table t;
    t.name = name;
    t.columns = cols;
    t.constraint_degree = degree;
    t.permutation_pairs = std::move(pp);
    t.eval_base = eb;
    t.eval_ext = ee;
    return t;
}
#define ORC_TABLE(name, ns, cols, degree, ...) make_table(name, cols, degree, ns::eval<FOps>, ns::eval<EOps>, ##__VA_ARGS__)

inline bool table_available(int id) { return id == T_CPU || id == T_MEMORY || id == T_CMP || id == T_RANGECHECK; }

inline Table table_by_id(int id) {
    switch (id) {
        case T_CPU: return ORC_TABLE("CpuStark", cpu_t, ola::air::cpu::NUM_CPU_COLS, 7);
        case T_MEMORY: return ORC_TABLE("MemoryStark", mem_t, ola::air::mem::NUM_MEM_COLS, 8);
        case T_CMP: return ORC_TABLE("CmpStark", cmp, cmp::NUM, 3);
        case T_RANGECHECK:
            return ORC_TABLE("RangeCheckStark", rangecheck, rangecheck::NUM, 3,
                             {PermutationPair{{{rangecheck::LIMB_LO, rangecheck::LIMB_LO_PERMUTED}}},
                              PermutationPair{{{rangecheck::LIMB_HI, rangecheck::LIMB_HI_PERMUTED}}},
                              PermutationPair{{{rangecheck::FIX_RANGE_CHECK_U16, rangecheck::FIX_RANGE_CHECK_U16_PERMUTED_LO}}},
                              PermutationPair{{{rangecheck::FIX_RANGE_CHECK_U16, rangecheck::FIX_RANGE_CHECK_U16_PERMUTED_HI}}}});
        default: throw std::runtime_error("table not restated yet");
    }
}

/* all_cross_table_lookups (ola_stark.rs:121-143) via the shared registry */
struct RegPolicy {
    typedef orc::Column Column;
    typedef TableWithColumns Twc;
    typedef CrossTableLookup Ctl;
    static Column single(int c) { return Column::single(c); }
    static Column linear(std::vector<std::pair<int, uint64_t>> v, uint64_t k) { return Column::linear(std::vector<std::pair<int, F>>(v.begin(), v.end()), k); }
    static Twc twc(int table, std::vector<Column> cols, Column filter) { return orc::twc(table, std::move(cols), std::move(filter)); }
};
inline std::vector<CrossTableLookup> all_cross_table_lookups() { return ola::air::build_ctl_registry<RegPolicy>(); }

/* A proving system = an ordered subset of the 12 tables (proof order = enum order) plus every registered CTL side
 * whose table is inside the subset (table ids remapped to positions).  CTLs losing a side become partial. */
inline System make_system(const std::vector<int>& ids) {
    System s;
    std::vector<int> pos(T_NUM, -1);
    for (size_t i = 0; i < ids.size(); i++) { pos[ids[i]] = (int)i; s.tables.push_back(table_by_id(ids[i])); }
    for (auto ctl : all_cross_table_lookups()) {
        CrossTableLookup out;
        out.complete = ctl.has_looked && !ctl.missing_sides;
        for (auto& l : ctl.looking) {
            if (pos[l.table] >= 0) { l.table = pos[l.table]; out.looking.push_back(l); }
            else out.complete = false;
        }
        out.has_looked = ctl.has_looked && pos[ctl.looked.table] >= 0;
        out.complete = out.complete && out.has_looked;
        if (out.has_looked) { out.looked = ctl.looked; out.looked.table = pos[ctl.looked.table]; }
        if (out.looking.empty() && !out.has_looked) continue;
        s.ctls.push_back(out);
    }
    s.compress_challenges.assign(ids.size(), 0);
    return s;
}

}  // namespace orc
#endif
