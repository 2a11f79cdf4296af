/* ORACLE (test infrastructure, NOT product code) -- CPU restatement of the AIR tables ("chips") and of the
 * cross-table-lookup registry.  Each eval<> follows the Rust eval_packed_generic of the cited file line by
 * line and emits the constraints in source order (the consumer is Horner in alpha, so order is semantics:
 * circuits/src/stark/constraint_consumer.rs:60-65).
 *
 * Independently restated here (enum order of circuits/src/stark/ola_stark.rs:104-119):
 *   3 Cmp         circuits/src/builtins/cmp/{columns.rs:16-22, cmp_stark.rs:21-45, :88-108}
 *   4 RangeCheck  circuits/src/builtins/rangecheck/{columns.rs:25-39, rangecheck_stark.rs:27-108, :111-140},
 *                 circuits/src/stark/lookup.rs:13-35
 */
#ifndef ORC_TABLES_HPP
#define ORC_TABLES_HPP
#include "stark.hpp"
#include "poseidon_constants.h"
/* every table's AIR and the CTL registry are restated in this directory from the Rust (air_*.hpp, ctl_registry.hpp); nothing
 * under olavm_b200/ is included */
#include "ctl_registry.hpp"

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

/* ---- the other ten tables: air_cpu.hpp, air_memory.hpp, air_builtins.hpp, air_storage_program.hpp ---- */
namespace cpu_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { cpu_air::eval<O>(lv, nv, yc); } }
namespace mem_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { memory_air::eval<O>(lv, nv, yc); } }
namespace tape_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { tape_air::eval<O>(lv, nv, yc); } }
namespace sccall_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { sccall_air::eval<O>(lv, nv, yc); } }
namespace prog_chunk_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { prog_chunk_air::eval<O>(lv, nv, yc); } }
namespace storage_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { storage_air::eval<O>(lv, nv, yc); } }
namespace psdn_chunk_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { poseidon_chunk_air::eval<O>(lv, nv, yc); } }
namespace psdn_t { template <class O> void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { poseidon_air::eval<O>(lv, nv, yc); } }

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

inline bool table_available(int id) { return id >= 0 && id < T_NUM; }

/* beta: the table's compress challenge (Bitwise / Program only; verifier.rs:78-86 takes it from the proof) */
inline Table table_by_id(int id, F beta = 0) {
    switch (id) {
        case T_CPU: return ORC_TABLE("CpuStark", cpu_t, cpu_air::NUM_COLS, 7);
        case T_MEMORY: return ORC_TABLE("MemoryStark", mem_t, memory_air::NUM_COLS, 8);
        case T_CMP: return ORC_TABLE("CmpStark", cmp, cmp::NUM, 3);
        case T_RANGECHECK:
            return ORC_TABLE("RangeCheckStark", rangecheck, rangecheck::NUM, 3,
                             {PermutationPair{{{rangecheck::LIMB_LO, rangecheck::LIMB_LO_PERMUTED}}},
                              PermutationPair{{{rangecheck::LIMB_HI, rangecheck::LIMB_HI_PERMUTED}}},
                              PermutationPair{{{rangecheck::FIX_RANGE_CHECK_U16, rangecheck::FIX_RANGE_CHECK_U16_PERMUTED_LO}}},
                              PermutationPair{{{rangecheck::FIX_RANGE_CHECK_U16, rangecheck::FIX_RANGE_CHECK_U16_PERMUTED_HI}}}});
        case T_POSEIDON: return ORC_TABLE("PoseidonStark", psdn_t, poseidon_air::NUM_COLS, 7);
        case T_POSEIDON_CHUNK: return ORC_TABLE("PoseidonChunkStark", psdn_chunk_t, poseidon_chunk_air::NUM_COLS, 3);
        case T_STORAGE: return ORC_TABLE("StorageAccessStark", storage_t, storage_air::NUM_COLS, 4);
        case T_TAPE: return ORC_TABLE("TapeStark", tape_t, tape_air::NUM_COLS, 5);
        case T_SCCALL: return ORC_TABLE("SCCallStark", sccall_t, sccall_air::NUM_COLS, 1);
        case T_PROG_CHUNK: return ORC_TABLE("ProgChunkStark", prog_chunk_t, prog_chunk_air::NUM_COLS, 4);
        case T_BITWISE: {
            namespace B = bitwise_air;
            std::vector<PermutationPair> pp; /* bitwise_stark.rs:352-363 */
            for (int i = 0; i < 4; ++i) pp.push_back(PermutationPair{{{B::COMPRESS_LIMBS + i, B::COMPRESS_PERMUTED + i}}});
            for (int i = 0; i < 4; ++i) pp.push_back(PermutationPair{{{B::FIX_COMPRESS, B::FIX_COMPRESS_PERMUTED + i}}});
            return make_table("BitwiseStark", B::NUM_COLS, 3,
                              [beta](const P<FOps>* lv, const P<FOps>* nv, Consumer<FOps>& yc) { B::eval<FOps>(lv, nv, yc, P<FOps>::c(beta)); },
                              [beta](const P<EOps>* lv, const P<EOps>* nv, Consumer<EOps>& yc) { B::eval<EOps>(lv, nv, yc, P<EOps>::c(beta)); }, pp);
        }
        case T_PROGRAM: {
            namespace G = program_air;
            return make_table("ProgramStark", G::NUM_COLS, 3,
                              [beta](const P<FOps>* lv, const P<FOps>* nv, Consumer<FOps>& yc) { G::eval<FOps>(lv, nv, yc, P<FOps>::c(beta)); },
                              [beta](const P<EOps>* lv, const P<EOps>* nv, Consumer<EOps>& yc) { G::eval<EOps>(lv, nv, yc, P<EOps>::c(beta)); },
                              {PermutationPair{{{G::COMP_PROG, G::COMP_PROG_PERM}}}, PermutationPair{{{G::EXEC_COMP_PROG, G::EXEC_COMP_PROG_PERM}}}}); /* program_stark.rs:110-115 */
        }
        default: throw std::runtime_error("unknown table id");
    }
}

/* all_cross_table_lookups (ola_stark.rs:121-143): ctl_registry.hpp */
inline std::vector<CrossTableLookup> all_cross_table_lookups() { return ctl::all_cross_table_lookups(); }

/* A proving system = an ordered subset of the 12 tables (proof order = enum order) plus every registered CTL side
 * whose table is inside the subset (table ids remapped to positions).  CTLs losing a side become partial. */
inline System make_system(const std::vector<int>& ids, const VF& compress = VF()) {
    System s;
    std::vector<int> pos(T_NUM, -1);
    s.compress_challenges.assign(ids.size(), 0);
    for (size_t i = 0; i < ids.size(); i++) {
        pos[ids[i]] = (int)i;
        F beta = (i < compress.size() && (ids[i] == T_BITWISE || ids[i] == T_PROGRAM)) ? gl_canon(compress[i]) : 0;
        s.compress_challenges[i] = beta;
        s.tables.push_back(table_by_id(ids[i], beta));
    }
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
    return s;
}

}  // namespace orc
#endif
