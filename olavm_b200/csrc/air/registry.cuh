// Table registry: metadata (columns, degree, permutation pairs) of every table whose constraint kernel is compiled
// in, and the cross-table-lookup registry of circuits/src/stark/ola_stark.rs:121-642 restricted to those tables.
#pragma once
#include "../stark_types.h"
#include "builtins_small.cuh"

namespace ola {
namespace stark {

inline bool table_available(int id) { return id == T_CMP || id == T_RANGECHECK; }

inline TableInfo table_info(int id) {
    TableInfo t;
    t.id = id;
    switch (id) {
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
        default: throw Error(OLA_ERR_INVALID_ARG, "table " + std::to_string(id) + " has no constraint kernel in this build");
    }
    return t;
}

inline std::vector<CrossTableLookup> all_cross_table_lookups() {
    std::vector<CrossTableLookup> v;
    // ctl_cmp_rangecheck (ola_stark.rs:282-296): looking RangeCheck(VAL | CMP_FILTER), looked Cmp(abs_diff | filter_looking_rc)
    v.push_back({{twc(T_RANGECHECK, singles({air::RangeCheck::VAL}), Column::single(air::RangeCheck::CMP_FILTER))},
                 twc(T_CMP, singles({air::Cmp::ABS_DIFF}), Column::single(air::Cmp::FILTER_LOOKING_RC))});
    return v;
}

// An ordered subset of the 12 tables (proof order = enum order) plus every registered CTL inside the subset.
inline System make_system(const std::vector<int>& ids) {
    System s;
    std::vector<int> pos(T_NUM, -1);
    for (size_t i = 0; i < ids.size(); i++) {
        OLA_CHECK(ids[i] >= 0 && ids[i] < T_NUM && pos[ids[i]] < 0, OLA_ERR_INVALID_ARG, "bad or duplicate table id");
        pos[ids[i]] = (int)i;
        s.tables.push_back(table_info(ids[i]));
    }
    for (auto ctl : all_cross_table_lookups()) {
        bool ok = pos[ctl.looked.table] >= 0;
        for (auto& l : ctl.looking) ok = ok && pos[l.table] >= 0;
        if (!ok) continue;
        for (auto& l : ctl.looking) l.table = pos[l.table];
        ctl.looked.table = pos[ctl.looked.table];
        s.ctls.push_back(ctl);
    }
    s.compress_challenges.assign(ids.size(), 0);
    return s;
}

}  // namespace stark
}  // namespace ola
