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

/* ---- shared transcriptions of builtins_air.h / hash_air.h ---- */
#define ORC_SHARED_TABLE(NS_T, NS)                                                                     \
    namespace NS_T {                                                                                   \
    template <class O>                                                                                 \
    void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { ola::air::NS::eval<P<O>, const P<O>*, Consumer<O>>(lv, nv, yc); } \
    }
ORC_SHARED_TABLE(tape_t, tape)
ORC_SHARED_TABLE(sccall_t, sccall)
ORC_SHARED_TABLE(prog_chunk_t, prog_chunk)
ORC_SHARED_TABLE(storage_t, storage)
ORC_SHARED_TABLE(psdn_chunk_t, psdn_chunk)
#undef ORC_SHARED_TABLE
/* Poseidon table: parameter tables from the oracle's own generated constants */
struct PoseidonParams {
    static uint64_t round(int i) { return ORC_ALL_ROUND_CONSTANTS[i]; }
    static uint64_t circ(int i) { return ORC_MDS_MATRIX_CIRC[i]; }
    static uint64_t diag(int i) { return ORC_MDS_MATRIX_DIAG[i]; }
    static uint64_t first(int i) { return ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]; }
    static uint64_t partial(int r) { return ORC_FAST_PARTIAL_ROUND_CONSTANTS[r]; }
    static uint64_t init(int r, int c) { return ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r][c]; }
    static uint64_t what(int r, int i) { return ORC_FAST_PARTIAL_ROUND_W_HATS[r][i]; }
    static uint64_t vs(int r, int i) { return ORC_FAST_PARTIAL_ROUND_VS[r][i]; }
};
namespace psdn_t {
template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) { ola::air::psdn::eval<P<O>, const P<O>*, Consumer<O>, PoseidonParams>(lv, nv, yc); }
}  // namespace psdn_t

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
        case T_CPU: return ORC_TABLE("CpuStark", cpu_t, ola::air::cpu::NUM_CPU_COLS, 7);
        case T_MEMORY: return ORC_TABLE("MemoryStark", mem_t, ola::air::mem::NUM_MEM_COLS, 8);
        case T_CMP: return ORC_TABLE("CmpStark", cmp, cmp::NUM, 3);
        case T_RANGECHECK:
            return ORC_TABLE("RangeCheckStark", rangecheck, rangecheck::NUM, 3,
                             {PermutationPair{{{rangecheck::LIMB_LO, rangecheck::LIMB_LO_PERMUTED}}},
                              PermutationPair{{{rangecheck::LIMB_HI, rangecheck::LIMB_HI_PERMUTED}}},
                              PermutationPair{{{rangecheck::FIX_RANGE_CHECK_U16, rangecheck::FIX_RANGE_CHECK_U16_PERMUTED_LO}}},
                              PermutationPair{{{rangecheck::FIX_RANGE_CHECK_U16, rangecheck::FIX_RANGE_CHECK_U16_PERMUTED_HI}}}});
        case T_POSEIDON: return ORC_TABLE("PoseidonStark", psdn_t, ola::air::psdn::NUM_POSEIDON_COLS, 7);
        case T_POSEIDON_CHUNK: return ORC_TABLE("PoseidonChunkStark", psdn_chunk_t, ola::air::psdn_chunk::NUM_POSEIDON_CHUNK_COLS, 3);
        case T_STORAGE: return ORC_TABLE("StorageAccessStark", storage_t, ola::air::storage::NUM_COL_ST, 4);
        case T_TAPE: return ORC_TABLE("TapeStark", tape_t, ola::air::tape::NUM_COL_TAPE, 5);
        case T_SCCALL: return ORC_TABLE("SCCallStark", sccall_t, ola::air::sccall::NUM_COL_SCCALL, 1);
        case T_PROG_CHUNK: return ORC_TABLE("ProgChunkStark", prog_chunk_t, ola::air::prog_chunk::NUM_PROG_CHUNK_COLS, 4);
        case T_BITWISE: {
            namespace B = ola::air::bitwise;
            std::vector<PermutationPair> pp;
            for (int i = 0; i < 4; ++i) pp.push_back(PermutationPair{{{B::COMPRESS_LIMBS + i, B::COMPRESS_PERMUTED + i}}});
            for (int i = 0; i < 4; ++i) pp.push_back(PermutationPair{{{B::FIX_COMPRESS, B::FIX_COMPRESS_PERMUTED + i}}});
            return make_table("BitwiseStark", B::COL_NUM_BITWISE, 3,
                              [beta](const P<FOps>* lv, const P<FOps>* nv, Consumer<FOps>& yc) { B::eval<P<FOps>, const P<FOps>*, Consumer<FOps>>(lv, nv, yc, P<FOps>::c(beta)); },
                              [beta](const P<EOps>* lv, const P<EOps>* nv, Consumer<EOps>& yc) { B::eval<P<EOps>, const P<EOps>*, Consumer<EOps>>(lv, nv, yc, P<EOps>::c(beta)); }, pp);
        }
        case T_PROGRAM: {
            namespace G = ola::air::program;
            return make_table("ProgramStark", G::NUM_PROG_COLS, 3,
                              [beta](const P<FOps>* lv, const P<FOps>* nv, Consumer<FOps>& yc) { G::eval<P<FOps>, const P<FOps>*, Consumer<FOps>>(lv, nv, yc, P<FOps>::c(beta)); },
                              [beta](const P<EOps>* lv, const P<EOps>* nv, Consumer<EOps>& yc) { G::eval<P<EOps>, const P<EOps>*, Consumer<EOps>>(lv, nv, yc, P<EOps>::c(beta)); },
                              {PermutationPair{{{G::COL_PROG_COMP_PROG, G::COL_PROG_COMP_PROG_PERM}}}, PermutationPair{{{G::COL_PROG_EXEC_COMP_PROG, G::COL_PROG_EXEC_COMP_PROG_PERM}}}});
        }
        default: throw std::runtime_error("unknown table id");
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
