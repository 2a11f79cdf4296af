// Table registry: metadata (columns, degree, permutation pairs) of the 12 tables of circuits/src/stark/ola_stark.rs:27-41
// and the cross-table-lookup registry of ola_stark.rs:121-642.
#pragma once
#include "../stark_types.h"
#include "builtins_small.cuh"
#include "ctl_registry.h"

namespace ola {
namespace stark {

inline bool table_available(int id) { return id >= 0 && id < T_NUM; }

inline TableInfo table_info(int id) {
    TableInfo t;
    t.id = id;
    switch (id) {
        case T_CPU:
            t.name = "CpuStark";
            t.columns = air::Cpu::COLUMNS;
            t.constraint_degree = air::Cpu::CONSTRAINT_DEGREE;
            break;
        case T_MEMORY:
            t.name = "MemoryStark";
            t.columns = air::Memory::COLUMNS;
            t.constraint_degree = air::Memory::CONSTRAINT_DEGREE;
            break;
        case T_CMP:
            t.name = "CmpStark";
            t.columns = air::Cmp::COLUMNS;
            t.constraint_degree = air::Cmp::CONSTRAINT_DEGREE;
            break;
        case T_RANGECHECK: {
            typedef air::RangeCheck R;
            t.name = "RangeCheckStark";
            t.columns = R::COLUMNS;
            t.constraint_degree = R::CONSTRAINT_DEGREE;
            // rangecheck_stark.rs:100-107
            t.permutation_pairs = {PermutationPair{{{R::LIMB_LO, R::LIMB_LO_PERMUTED}}}, PermutationPair{{{R::LIMB_HI, R::LIMB_HI_PERMUTED}}},
                                   PermutationPair{{{R::FIX_RANGE_CHECK_U16, R::FIX_RANGE_CHECK_U16_PERMUTED_LO}}},
                                   PermutationPair{{{R::FIX_RANGE_CHECK_U16, R::FIX_RANGE_CHECK_U16_PERMUTED_HI}}}};
            break;
        }
#define OLA_TABLE_CASE(ID, NAME, AIR)                          \
    case ID:                                                  \
        t.name = NAME;                                        \
        t.columns = air::AIR::COLUMNS;                        \
        t.constraint_degree = air::AIR::CONSTRAINT_DEGREE;    \
        break;
            OLA_TABLE_CASE(T_POSEIDON, "PoseidonStark", Poseidon)
            OLA_TABLE_CASE(T_POSEIDON_CHUNK, "PoseidonChunkStark", PoseidonChunk)
            OLA_TABLE_CASE(T_STORAGE, "StorageAccessStark", StorageAccess)
            OLA_TABLE_CASE(T_TAPE, "TapeStark", Tape)
            OLA_TABLE_CASE(T_SCCALL, "SCCallStark", SCCall)
            OLA_TABLE_CASE(T_PROG_CHUNK, "ProgChunkStark", ProgChunk)
#undef OLA_TABLE_CASE
        case T_BITWISE: {
            namespace B = air::bitwise;
            t.name = "BitwiseStark";
            t.columns = air::Bitwise::COLUMNS;
            t.constraint_degree = air::Bitwise::CONSTRAINT_DEGREE;
            // bitwise_stark.rs:352-363 (only the compress lookups carry a permutation argument)
            for (int i = 0; i < 4; ++i) t.permutation_pairs.push_back(PermutationPair{{{B::COMPRESS_LIMBS + i, B::COMPRESS_PERMUTED + i}}});
            for (int i = 0; i < 4; ++i) t.permutation_pairs.push_back(PermutationPair{{{B::FIX_COMPRESS, B::FIX_COMPRESS_PERMUTED + i}}});
            break;
        }
        case T_PROGRAM: {
            namespace G = air::program;
            t.name = "ProgramStark";
            t.columns = air::Program::COLUMNS;
            t.constraint_degree = air::Program::CONSTRAINT_DEGREE;
            // program_stark.rs:110-115
            t.permutation_pairs = {PermutationPair{{{G::COL_PROG_COMP_PROG, G::COL_PROG_COMP_PROG_PERM}}},
                                   PermutationPair{{{G::COL_PROG_EXEC_COMP_PROG, G::COL_PROG_EXEC_COMP_PROG_PERM}}}};
            break;
        }
        default: throw Error(OLA_ERR_INVALID_ARG, "unknown table id " + std::to_string(id));
    }
    return t;
}

struct RegPolicy {
    typedef stark::Column Column;
    typedef TableWithColumns Twc;
    typedef CrossTableLookup Ctl;
    static Column single(int c) { return Column::single(c); }
    static Column linear(std::vector<std::pair<int, uint64_t>> v, uint64_t k) { return Column::linear(std::move(v), k); }
    static Twc twc(int table, std::vector<Column> cols, Column filter) { return stark::twc(table, std::move(cols), std::move(filter)); }
};
inline std::vector<CrossTableLookup> all_cross_table_lookups() { return air::build_ctl_registry<RegPolicy>(); }

// An ordered subset of the 12 tables (proof order = enum order) plus every registered CTL side whose table is in
// the subset.  A CTL that loses a side is "partial" (complete = false): its Z columns are still proven.
inline System make_system(const std::vector<int>& ids) {
    System s;
    std::vector<int> pos(T_NUM, -1);
    for (size_t i = 0; i < ids.size(); i++) {
        OLA_CHECK(ids[i] >= 0 && ids[i] < T_NUM && pos[ids[i]] < 0, OLA_ERR_INVALID_ARG, "bad or duplicate table id");
        pos[ids[i]] = (int)i;
        s.tables.push_back(table_info(ids[i]));
    }
    for (auto ctl : all_cross_table_lookups()) {
        CrossTableLookup out;
        out.complete = ctl.has_looked && !ctl.missing_sides;
        for (auto& l : ctl.looking) {
            if (pos[l.table] >= 0) {
                l.table = pos[l.table];
                out.looking.push_back(l);
            } else {
                out.complete = false;
            }
        }
        out.has_looked = ctl.has_looked && pos[ctl.looked.table] >= 0;
        out.complete = out.complete && out.has_looked;
        if (out.has_looked) {
            out.looked = ctl.looked;
            out.looked.table = pos[ctl.looked.table];
        }
        if (out.looking.empty() && !out.has_looked) continue;
        s.ctls.push_back(out);
    }
    s.compress_challenges.assign(ids.size(), 0);
    return s;
}

}  // namespace stark
}  // namespace ola
