/* ORACLE (test infrastructure, NOT product code) -- the cross-table-lookup registry, restated from the Rust on its own
 * (nothing under olavm_b200/ is included): all_cross_table_lookups and the ctl_* builders of
 * circuits/src/stark/ola_stark.rs:121-560, with the per-table ctl_data_* / ctl_filter_* functions they call:
 *   cpu/cpu_stark.rs:20-327, memory/memory_stark.rs:17-80, builtins/bitwise/bitwise_stark.rs, builtins/cmp/cmp_stark.rs,
 *   builtins/rangecheck/rangecheck_stark.rs, builtins/poseidon/{poseidon_stark,poseidon_chunk_stark}.rs,
 *   builtins/storage/storage_access_stark.rs, builtins/tape/tape_stark.rs, builtins/sccall/sccall_stark.rs,
 *   program/{program_stark,prog_chunk_stark}.rs.
 * Column::zero() / one() are constant columns; F::NEG_ONE is p - 1. */
#ifndef ORC_CTL_REGISTRY_HPP
#define ORC_CTL_REGISTRY_HPP
#include "air_cpu.hpp"
#include "air_memory.hpp"
#include "air_storage_program.hpp"

namespace orc {
namespace ctl {

enum { CPU = 0, MEMORY, BITWISE, CMP, RANGECHECK, POSEIDON, POSEIDON_CHUNK, STORAGE, TAPE, SCCALL, PROGRAM, PROG_CHUNK };
/* Cmp / RangeCheck column indices (builtins/cmp/columns.rs:16-22, builtins/rangecheck/columns.rs:25-39) */
enum { CMP_OP0 = 0, CMP_OP1, CMP_GTE, CMP_ABS_DIFF, CMP_ABS_DIFF_INV, CMP_FILTER_LOOKING_RC };
enum { RC_CPU_FILTER = 0, RC_MEMORY_SORT_FILTER, RC_MEMORY_REGION_FILTER, RC_CMP_FILTER, RC_VAL };

typedef std::vector<Column> Cols;
inline Column neg_plus_one(int col) { return Column::linear({{col, GL_P - 1}}, 1); } /* 1 - col */
inline Column col_plus(int col, uint64_t k) { return Column::linear({{col, 1}}, k); }
inline Column zero() { return Column::constant_(0); }
inline Column one() { return Column::constant_(1); }
inline Cols range(int start, int count) { Cols r; for (int i = 0; i < count; i++) r.push_back(Column::single(start + i)); return r; }
inline Cols cat(Cols a, const Cols& b) { a.insert(a.end(), b.begin(), b.end()); return a; }
inline CrossTableLookup make(std::vector<TableWithColumns> looking, TableWithColumns looked) {
    CrossTableLookup c;
    c.looking = std::move(looking);
    c.looked = std::move(looked);
    return c;
}

inline CrossTableLookup ctl_cpu_memory() {
    namespace C = cpu_air;
    namespace M = memory_air;
    std::vector<TableWithColumns> lookers;
    const Column f_store_load = Column::sum({C::S_MSTORE, C::S_MLOAD}), f_call_ret = Column::sum({C::S_CALL, C::S_RET});
    lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::AUX1, C::DST}), f_store_load));
    lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::OP0, C::DST}), f_call_ret));
    lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::AUX0, C::AUX1}), f_call_ret));
    lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::AUX0, C::AUX1}), Column::single(C::FILTER_TAPE_LOOKING)));
    const int sccall_addr[4] = {C::OP0, C::DST, C::AUX0, C::AUX1};
    for (int i = 0; i < 4; i++)
        lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, sccall_addr[i], C::ADDR_CODE + i}), Column::single(C::IS_SCCALL_EXT_LINE)));
    for (int i = 0; i < 4; i++)
        lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::S_OP0 + i, C::S_OP0 + 4 + i}), Column::single(C::IS_STORAGE_EXT_LINE)));
    for (int i = 0; i < 4; i++)
        lookers.push_back(twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::S_OP1 + i, C::S_OP1 + 4 + i}), Column::single(C::IS_STORAGE_EXT_LINE)));
    const Column mem_filter = Column::sum({M::S_MLOAD, M::S_MSTORE, M::S_CALL, M::S_RET, M::S_TLOAD, M::S_TSTORE, M::S_SCCALL, M::S_SSTORE, M::S_SLOAD});
    return make(lookers, twc(MEMORY, singles({M::TX_IDX, M::ENV_IDX, M::CLK, M::OP, M::ADDR, M::VALUE}), mem_filter));
}
inline CrossTableLookup ctl_memory_rc_sort() {
    namespace M = memory_air;
    return make({twc(MEMORY, singles({M::RC_VALUE}), Column::single(M::FILTER_LOOKING_RC))},
                twc(RANGECHECK, singles({RC_VAL}), Column::single(RC_MEMORY_SORT_FILTER)));
}
inline CrossTableLookup ctl_memory_rc_region() {
    namespace M = memory_air;
    return make({twc(MEMORY, singles({M::DIFF_ADDR_COND}), Column::single(M::FILTER_LOOKING_RC_COND))},
                twc(RANGECHECK, singles({RC_VAL}), Column::single(RC_MEMORY_REGION_FILTER)));
}
inline CrossTableLookup ctl_bitwise_cpu() {
    namespace C = cpu_air;
    namespace B = bitwise_air;
    return make({twc(CPU, singles({C::OPCODE, C::OP0, C::OP1, C::DST}), Column::single(C::S_BITWISE))},
                twc(BITWISE, singles({B::TAG, B::OP0, B::OP1, B::RES}), Column::single(B::FILTER)));
}
inline CrossTableLookup ctl_cmp_cpu() {
    namespace C = cpu_air;
    return make({twc(CPU, singles({C::OP0, C::OP1, C::DST}), Column::single(C::S_GTE))},
                twc(CMP, singles({CMP_OP0, CMP_OP1, CMP_GTE}), Column::single(CMP_FILTER_LOOKING_RC)));
}
inline CrossTableLookup ctl_cmp_rangecheck() {
    return make({twc(RANGECHECK, singles({RC_VAL}), Column::single(RC_CMP_FILTER))},
                twc(CMP, singles({CMP_ABS_DIFF}), Column::single(CMP_FILTER_LOOKING_RC)));
}
inline CrossTableLookup ctl_rangecheck_cpu() {
    namespace C = cpu_air;
    return make({twc(CPU, singles({C::OP1}), Column::single(C::S_RC))}, twc(RANGECHECK, singles({RC_VAL}), Column::single(RC_CPU_FILTER)));
}
inline CrossTableLookup ctl_cpu_poseidon_chunk() {
    namespace C = cpu_air;
    namespace K = poseidon_chunk_air;
    return make({twc(CPU, singles({C::TX_IDX, C::ENV_IDX, C::CLK, C::OPCODE, C::OP0, C::OP1, C::DST}), Column::single(C::S_PSDN))},
                twc(POSEIDON_CHUNK, singles({K::TX_IDX, K::ENV_IDX, K::CLK, K::OPCODE, K::OP0, K::OP1, K::DST}), Column::single(K::FILTER_LOOKED_CPU)));
}
inline CrossTableLookup ctl_poseidon_chunk_mem() {
    namespace K = poseidon_chunk_air;
    namespace M = memory_air;
    std::vector<TableWithColumns> lookers;
    for (int i = 0; i < 8; i++) {
        Cols c = singles({K::TX_IDX, K::ENV_IDX, K::CLK, K::OPCODE});
        c.push_back(col_plus(K::OP0, (uint64_t)i));
        c.push_back(Column::single(K::VALUE + i));
        c.push_back(zero());
        lookers.push_back(twc(POSEIDON_CHUNK, c, Column::single(K::FILTER_LOOKING_MEM + i)));
    }
    for (int i = 0; i < 4; i++) {
        Cols c = singles({K::TX_IDX, K::ENV_IDX, K::CLK, K::OPCODE});
        c.push_back(col_plus(K::DST, (uint64_t)i));
        c.push_back(Column::single(K::HASH + i));
        c.push_back(one());
        lookers.push_back(twc(POSEIDON_CHUNK, c, Column::single(K::IS_RESULT_LINE)));
    }
    return make(lookers, twc(MEMORY, singles({M::TX_IDX, M::ENV_IDX, M::CLK, M::OP, M::ADDR, M::VALUE, M::IS_WRITE}), Column::single(M::S_POSEIDON)));
}
inline CrossTableLookup ctl_chunk_poseidon() {
    namespace K = poseidon_chunk_air;
    namespace G = prog_chunk_air;
    namespace H = poseidon_air;
    return make({twc(POSEIDON_CHUNK, cat(cat(range(K::VALUE, 8), range(K::CAP, 4)), range(K::HASH, 12)), Column::single(K::FILTER_LOOKING_POSEIDON)),
                 twc(PROG_CHUNK, cat(cat(range(G::INST, 8), range(G::CAP, 4)), range(G::HASH, 12)), neg_plus_one(G::IS_PADDING_LINE))},
                twc(POSEIDON, cat(range(H::INPUT, 12), range(H::OUTPUT, 12)), Column::single(H::FILTER_LOOKED_NORMAL)));
}
inline CrossTableLookup ctl_cpu_poseidon_tree_key() {
    namespace C = cpu_air;
    namespace H = poseidon_air;
    Cols c = cat(range(C::ADDR_STORAGE, 4), range(C::S_OP0 + 4, 4));
    for (int i = 0; i < 4; i++) c.push_back(zero());
    c = cat(c, range(C::S_DST, 4));
    return make({twc(CPU, c, Column::single(C::IS_STORAGE_EXT_LINE))},
                twc(POSEIDON, cat(range(H::INPUT, 12), range(H::OUTPUT, 4)), Column::single(H::FILTER_LOOKED_TREEKEY)));
}
inline CrossTableLookup ctl_cpu_storage_access() {
    namespace C = cpu_air;
    namespace S = storage_air;
    const Cols cpu = singles({C::IDX_STORAGE, C::S_SSTORE, C::S_DST, C::S_DST + 1, C::S_DST + 2, C::S_DST + 3, C::S_OP1 + 4, C::S_OP1 + 5, C::S_OP1 + 6, C::S_OP1 + 7});
    const Cols st = cat(singles({S::ACCESS_IDX, S::IS_WRITE}), cat(range(S::ADDR, 4), range(S::PATH, 4)));
    return make({twc(CPU, cpu, Column::single(C::IS_STORAGE_EXT_LINE))},
                twc(STORAGE, st, Column::linear({{S::IS_LAYER_256, 1}, {S::FILTER_IS_FOR_PROG, GL_P - 1}}, 0)));
}
inline CrossTableLookup ctl_storage_access_poseidon() {
    namespace S = storage_air;
    namespace H = poseidon_air;
    auto side = [](int first, int second, int hash) {
        Cols c = cat(range(first, 4), range(second, 4));
        c.push_back(Column::single(S::HASH_TYPE));
        for (int i = 0; i < 3; i++) c.push_back(zero());
        c = cat(c, range(hash, 4));
        c.push_back(Column::single(S::IS_LAYER_256));
        c.push_back(neg_plus_one(S::IS_LAYER_256));
        return c;
    };
    const Column bit0 = Column::single(S::FILTER_IS_HASH_BIT_0), bit1 = Column::single(S::FILTER_IS_HASH_BIT_1);
    std::vector<TableWithColumns> lookers = {twc(STORAGE, side(S::PATH, S::SIB, S::HASH), bit0), twc(STORAGE, side(S::PRE_PATH, S::SIB, S::PRE_HASH), bit0),
                                             twc(STORAGE, side(S::SIB, S::PATH, S::HASH), bit1), twc(STORAGE, side(S::SIB, S::PRE_PATH, S::PRE_HASH), bit1)};
    const Cols looked = cat(cat(range(H::INPUT, 12), range(H::OUTPUT, 4)), singles({H::FILTER_LOOKED_STORAGE_LEAF, H::FILTER_LOOKED_STORAGE_BRANCH}));
    return make(lookers, twc(POSEIDON, looked, Column::sum({H::FILTER_LOOKED_STORAGE_LEAF, H::FILTER_LOOKED_STORAGE_BRANCH})));
}
inline CrossTableLookup ctl_cpu_tape() {
    namespace C = cpu_air;
    namespace T = tape_air;
    std::vector<TableWithColumns> lookers;
    lookers.push_back(twc(CPU, singles({C::TX_IDX, C::OPCODE, C::S_OP0, C::AUX1}), Column::single(C::FILTER_TAPE_LOOKING)));
    const int value_base[3] = {C::S_OP0, C::ADDR_CODE, C::ADDR_STORAGE}; /* caller, callee code, callee storage: tp + 0.., 4.., 8.. */
    for (int g = 0; g < 3; g++)
        for (int i = 0; i < 4; i++) {
            Cols c = singles({C::TX_IDX, C::OPCODE});
            c.push_back(col_plus(C::TP, (uint64_t)(4 * g + i)));
            c.push_back(Column::single(value_base[g] + i));
            lookers.push_back(twc(CPU, c, Column::single(C::IS_SCCALL_EXT_LINE)));
        }
    return make(lookers, twc(TAPE, singles({T::TX_IDX, T::OPCODE, T::ADDR, T::VALUE}), Column::single(T::FILTER_LOOKED)));
}
inline CrossTableLookup ctl_cpu_sccall() {
    namespace C = cpu_air;
    namespace S = sccall_air;
    Cols cpu = singles({C::TX_IDX, C::ENV_IDX});
    cpu = cat(cpu, range(C::S_OP0, 8));
    cpu = cat(cpu, singles({C::CLK, C::OP1_IMM}));
    cpu = cat(cpu, range(C::REGS, C::REGISTER_NUM));
    cpu.push_back(col_plus(C::ENV_IDX, 1));
    Cols sc = singles({S::TX_IDX, S::CALLER_ENV_IDX});
    sc = cat(sc, range(S::CALLER_EXE_CTX, 4));
    sc = cat(sc, range(S::CALLER_CODE_CTX, 4));
    sc = cat(sc, singles({S::CLK_CALLER_CALL, S::CALLER_OP1_IMM}));
    sc = cat(sc, range(S::CALLER_REG, 10));
    sc.push_back(Column::single(S::CALLEE_ENV_IDX));
    return make({twc(CPU, cpu, Column::single(C::IS_SCCALL_EXT_LINE))}, twc(SCCALL, sc, neg_plus_one(S::IS_PADDING)));
}
inline CrossTableLookup ctl_cpu_sccall_end() {
    namespace C = cpu_air;
    namespace S = sccall_air;
    Cols cpu = singles({C::TX_IDX, C::ENV_IDX});
    cpu = cat(cpu, range(C::ADDR_STORAGE, 4));
    cpu = cat(cpu, range(C::ADDR_CODE, 4));
    cpu.push_back(Column::single(C::CLK));
    cpu = cat(cpu, range(C::REGS, C::REGISTER_NUM));
    cpu = cat(cpu, singles({C::AUX0, C::AUX1}));
    Cols sc = singles({S::TX_IDX, S::CALLER_ENV_IDX});
    sc = cat(sc, range(S::CALLER_EXE_CTX, 4));
    sc = cat(sc, range(S::CALLER_CODE_CTX, 4));
    sc.push_back(Column::single(S::CLK_CALLER_CALL));
    sc = cat(sc, range(S::CALLER_REG, 10));
    sc = cat(sc, singles({S::CALLEE_ENV_IDX, S::CLK_CALLEE_END}));
    return make({twc(CPU, cpu, Column::single(C::FILTER_SCCALL_END))}, twc(SCCALL, sc, neg_plus_one(S::IS_PADDING)));
}
inline CrossTableLookup ctl_cpu_program() {
    namespace C = cpu_air;
    namespace G = program_air;
    Cols inst = cat(range(C::ADDR_CODE, 4), singles({C::PC, C::INST}));
    Cols imm = range(C::ADDR_CODE, 4);
    imm.push_back(col_plus(C::PC, 1));
    imm.push_back(Column::single(C::IMM_VAL));
    return make({twc(CPU, inst, Column::linear({{C::IS_EXT_LINE, GL_P - 1}, {C::IS_PADDING, GL_P - 1}}, 1)),
                 twc(CPU, imm, Column::single(C::FILTER_LOOKING_PROG_IMM))},
                twc(PROGRAM, cat(range(G::EXEC_CODE_ADDR, 4), singles({G::EXEC_PC, G::EXEC_INST})), Column::single(G::FILTER_EXEC)));
}
inline CrossTableLookup ctl_prog_chunk_prog() {
    namespace K = prog_chunk_air;
    namespace G = program_air;
    std::vector<TableWithColumns> lookers;
    for (int i = 0; i < 8; i++) {
        Cols c = range(K::CODE_ADDR, 4);
        c.push_back(col_plus(K::START_PC, (uint64_t)i));
        c.push_back(Column::single(K::INST + i));
        lookers.push_back(twc(PROG_CHUNK, c, Column::single(K::FILTER_LOOKING_PROG + i)));
    }
    return make(lookers, twc(PROGRAM, cat(range(G::CODE_ADDR, 4), singles({G::PC, G::INST})), Column::single(G::FILTER_PROG_CHUNK)));
}
inline CrossTableLookup ctl_prog_chunk_storage() {
    namespace K = prog_chunk_air;
    namespace S = storage_air;
    Cols c = {zero()};
    c = cat(c, cat(range(K::CODE_ADDR, 4), range(K::HASH, 4)));
    return make({twc(PROG_CHUNK, c, Column::single(K::IS_RESULT_LINE))},
                twc(STORAGE, cat(singles({S::IS_WRITE}), cat(range(S::ADDR, 4), range(S::PATH, 4))), Column::single(S::FILTER_IS_FOR_PROG)));
}

/* all_cross_table_lookups (ola_stark.rs:121-143), registry order */
inline std::vector<CrossTableLookup> all_cross_table_lookups() {
    return {ctl_cpu_memory(),          ctl_memory_rc_sort(),       ctl_memory_rc_region(), ctl_bitwise_cpu(),      ctl_cmp_cpu(),
            ctl_cmp_rangecheck(),      ctl_rangecheck_cpu(),       ctl_cpu_poseidon_chunk(), ctl_poseidon_chunk_mem(), ctl_chunk_poseidon(),
            ctl_cpu_poseidon_tree_key(), ctl_cpu_storage_access(), ctl_storage_access_poseidon(), ctl_cpu_tape(),  ctl_cpu_sccall(),
            ctl_cpu_sccall_end(),      ctl_cpu_program(),          ctl_prog_chunk_prog(),  ctl_prog_chunk_storage()};
}

}  // namespace ctl
}  // namespace orc
#endif
