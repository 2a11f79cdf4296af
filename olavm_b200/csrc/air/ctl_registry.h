// The cross-table-lookup registry (all_cross_table_lookups, circuits/src/stark/ola_stark.rs:121-143 and the
// ctl_* builders :146-642), in registry order, as DATA shared by the product (device descriptors) and the oracle.
// It is generic over a policy `Pol` that supplies each side's Column / TableWithColumns / CrossTableLookup types:
//   Pol::Column, Pol::Twc, Pol::Ctl{looking, looked, has_looked, missing_sides}, Pol::single(c), Pol::linear({(c,k)..}, const),
//   Pol::twc(table, columns, filter)
// All 19 entries carry every side; a CTL becomes "partial" only when a proving system is built over a subset of
// the 12 tables (make_system).
//
// Column definitions cited per entry: circuits/src/cpu/cpu_stark.rs:17-330 (CPU sides),
// builtins/cmp/cmp_stark.rs:88-108, builtins/rangecheck/rangecheck_stark.rs:111-150.
#pragma once
#include <utility>
#include <vector>

#include "cpu_air.h"
#include "hash_air.h"
#include "mem_air.h"

namespace ola {
namespace air {

enum RegTable { RT_CPU = 0, RT_MEMORY, RT_BITWISE, RT_CMP, RT_RANGECHECK, RT_POSEIDON, RT_POSEIDON_CHUNK, RT_STORAGE, RT_TAPE, RT_SCCALL, RT_PROGRAM, RT_PROG_CHUNK };

template <class Pol>
std::vector<typename Pol::Ctl> build_ctl_registry() {
    using namespace cpu;
    typedef typename Pol::Column Col;
    typedef typename Pol::Twc Twc;
    typedef typename Pol::Ctl Ctl;
    typedef std::vector<Col> Cols;
    const uint64_t NEG_ONE = 0xFFFFFFFF00000000ULL;
    auto S = [](std::initializer_list<int> cs) { Cols v; for (int c : cs) v.push_back(Pol::single(c)); return v; };
    auto sum = [](std::initializer_list<int> cs) { std::vector<std::pair<int, uint64_t>> v; for (int c : cs) v.push_back({c, 1}); return Pol::linear(v, 0); };
    auto plus = [](int c, uint64_t k) { return Pol::linear({{c, 1}}, k); };

    auto full = [](std::vector<Twc> looking, Twc looked) { Ctl c; c.looking = std::move(looking); c.looked = std::move(looked); c.has_looked = true; return c; };
    // rangecheck / cmp column ids (builtins/rangecheck/columns.rs:25-39, builtins/cmp/columns.rs:16-22)
    const int RC_CPU_FILTER = 0, RC_CMP_FILTER = 3, RC_VAL = 4;
    const int CMP_OP0 = 0, CMP_OP1 = 1, CMP_GTE = 2, CMP_ABS_DIFF = 3, CMP_FILTER = 5;

    std::vector<Ctl> v;
    // Memory sides: circuits/src/memory/memory_stark.rs:17-75
    using namespace mem;
    // 1. ctl_cpu_memory (:146-200): 16 CPU lookers -> Memory(tx, env, clk, op, addr, value | sum of 9 selectors)
    {
        std::vector<Twc> l;
        l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_AUX1, COL_DST}), sum({COL_S_MSTORE, COL_S_MLOAD})));
        l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_OP0, COL_DST}), sum({COL_S_CALL, COL_S_RET})));
        l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_AUX0, COL_AUX1}), sum({COL_S_CALL, COL_S_RET})));
        l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_AUX0, COL_AUX1}), Pol::single(COL_FILTER_TAPE_LOOKING)));
        const int sc_addr[4] = {COL_OP0, COL_DST, COL_AUX0, COL_AUX1};
        for (int i = 0; i < 4; ++i)
            l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, sc_addr[i], COL_ADDR_CODE + i}), Pol::single(IS_SCCALL_EXT_LINE)));
        for (int i = 0; i < 4; ++i)
            l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_S_OP0 + i, COL_S_OP0 + 4 + i}), Pol::single(COL_IS_STORAGE_EXT_LINE)));
        for (int i = 0; i < 4; ++i)
            l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_S_OP1 + i, COL_S_OP1 + 4 + i}), Pol::single(COL_IS_STORAGE_EXT_LINE)));
        v.push_back(full(l, Pol::twc(RT_MEMORY, S({COL_MEM_TX_IDX, COL_MEM_ENV_IDX, COL_MEM_CLK, COL_MEM_OP, COL_MEM_ADDR, COL_MEM_VALUE}),
                                     sum({COL_MEM_S_MLOAD, COL_MEM_S_MSTORE, COL_MEM_S_CALL, COL_MEM_S_RET, COL_MEM_S_TLOAD, COL_MEM_S_TSTORE, COL_MEM_S_SCCALL,
                                          COL_MEM_S_SSTORE, COL_MEM_S_SLOAD}))));
    }
    // 2. ctl_memory_rc_sort, 3. ctl_memory_rc_region (:202-231): Memory -> RangeCheck
    v.push_back(full({Pol::twc(RT_MEMORY, S({COL_MEM_RC_VALUE}), Pol::single(COL_MEM_FILTER_LOOKING_RC))},
                     Pol::twc(RT_RANGECHECK, S({RC_VAL}), Pol::single(1 /* MEMORY_SORT_FILTER */))));
    v.push_back(full({Pol::twc(RT_MEMORY, S({COL_MEM_DIFF_ADDR_COND}), Pol::single(COL_MEM_FILTER_LOOKING_RC_COND))},
                     Pol::twc(RT_RANGECHECK, S({RC_VAL}), Pol::single(2 /* MEMORY_REGION_FILTER */))));
    // 4. ctl_bitwise_cpu (:251-265): CPU -> Bitwise(tag, op0, op1, res | filter)  (bitwise_stark.rs:365-371)
    v.push_back(full({Pol::twc(RT_CPU, S({COL_OPCODE, COL_OP0, COL_OP1, COL_DST}), Pol::single(COL_S_BITWISE))},
                     Pol::twc(RT_BITWISE, S({bitwise::TAG, bitwise::OP0, bitwise::OP1, bitwise::RES}), Pol::single(bitwise::FILTER))));
    // 5. ctl_cmp_cpu (:268-281)
    v.push_back(full({Pol::twc(RT_CPU, S({COL_OP0, COL_OP1, COL_DST}), Pol::single(COL_S_GTE))},
                     Pol::twc(RT_CMP, S({CMP_OP0, CMP_OP1, CMP_GTE}), Pol::single(CMP_FILTER))));
    // 6. ctl_cmp_rangecheck (:283-296)
    v.push_back(full({Pol::twc(RT_RANGECHECK, S({RC_VAL}), Pol::single(RC_CMP_FILTER))}, Pol::twc(RT_CMP, S({CMP_ABS_DIFF}), Pol::single(CMP_FILTER))));
    // 7. ctl_rangecheck_cpu (:299-312)
    v.push_back(full({Pol::twc(RT_CPU, S({COL_OP1}), Pol::single(COL_S_RC))}, Pol::twc(RT_RANGECHECK, S({RC_VAL}), Pol::single(RC_CPU_FILTER))));
    // 8. ctl_cpu_poseidon_chunk (:314-328): CPU -> PoseidonChunk  (poseidon_chunk_stark.rs:23-39)
    {
        using namespace psdn_chunk;
        v.push_back(full({Pol::twc(RT_CPU, S({COL_TX_IDX, COL_ENV_IDX, COL_CLK, COL_OPCODE, COL_OP0, COL_OP1, COL_DST}), Pol::single(COL_S_PSDN))},
                         Pol::twc(RT_POSEIDON_CHUNK, S({COL_POSEIDON_CHUNK_TX_IDX, COL_POSEIDON_CHUNK_ENV_IDX, COL_POSEIDON_CHUNK_CLK, COL_POSEIDON_CHUNK_OPCODE,
                                                         COL_POSEIDON_CHUNK_OP0, COL_POSEIDON_CHUNK_OP1, COL_POSEIDON_CHUNK_DST}),
                                  Pol::single(COL_POSEIDON_CHUNK_FILTER_LOOKED_CPU))));
    }
    // 9. ctl_poseidon_chunk_mem (:330-356): 8 src + 4 dst PoseidonChunk lookers (poseidon_chunk_stark.rs:41-81)
    //    -> Memory(tx, env, clk, op, addr, value, is_write | s_poseidon)
    {
        using namespace psdn_chunk;
        std::vector<Twc> l;
        for (int i = 0; i < 8; ++i) {
            Cols c = S({COL_POSEIDON_CHUNK_TX_IDX, COL_POSEIDON_CHUNK_ENV_IDX, COL_POSEIDON_CHUNK_CLK, COL_POSEIDON_CHUNK_OPCODE});
            c.push_back(plus(COL_POSEIDON_CHUNK_OP0, (uint64_t)i));
            c.push_back(Pol::single(COL_POSEIDON_CHUNK_VALUE + i));
            c.push_back(Pol::linear({}, 0));  // Column::zero()
            l.push_back(Pol::twc(RT_POSEIDON_CHUNK, c, Pol::single(COL_POSEIDON_CHUNK_FILTER_LOOKING_MEM + i)));
        }
        for (int i = 0; i < 4; ++i) {
            Cols c = S({COL_POSEIDON_CHUNK_TX_IDX, COL_POSEIDON_CHUNK_ENV_IDX, COL_POSEIDON_CHUNK_CLK, COL_POSEIDON_CHUNK_OPCODE});
            c.push_back(plus(COL_POSEIDON_CHUNK_DST, (uint64_t)i));
            c.push_back(Pol::single(COL_POSEIDON_CHUNK_HASH + i));
            c.push_back(Pol::linear({}, 1));  // Column::one()
            l.push_back(Pol::twc(RT_POSEIDON_CHUNK, c, Pol::single(COL_POSEIDON_CHUNK_IS_RESULT_LINE)));
        }
        v.push_back(full(l, Pol::twc(RT_MEMORY, S({COL_MEM_TX_IDX, COL_MEM_ENV_IDX, COL_MEM_CLK, COL_MEM_OP, COL_MEM_ADDR, COL_MEM_VALUE, COL_MEM_IS_WRITE}),
                                     Pol::single(COL_MEM_S_POSEIDON))));
    }
    // 10. ctl_chunk_poseidon (:358-379): PoseidonChunk(value, cap, hash | filter_looking_poseidon) and
    //     ProgChunk(inst, cap, hash | 1 - is_padding) -> Poseidon(input, output | filter_looked_normal)
    //     (poseidon_chunk_stark.rs:83-95, prog_chunk_stark.rs:39-50, poseidon_stark.rs:159-165)
    {
        Cols pc, pg, ps;
        for (int i = 0; i < 24; ++i) pc.push_back(Pol::single(psdn_chunk::COL_POSEIDON_CHUNK_VALUE + i));
        for (int i = 0; i < 24; ++i) pg.push_back(Pol::single(prog_chunk::COL_PROG_CHUNK_INST + i));
        for (int i = 0; i < 24; ++i) ps.push_back(Pol::single(psdn::COL_POSEIDON_INPUT + i));
        v.push_back(full({Pol::twc(RT_POSEIDON_CHUNK, pc, Pol::single(psdn_chunk::COL_POSEIDON_CHUNK_FILTER_LOOKING_POSEIDON)),
                          Pol::twc(RT_PROG_CHUNK, pg, Pol::linear({{prog_chunk::COL_PROG_CHUNK_IS_PADDING_LINE, NEG_ONE}}, 1))},
                         Pol::twc(RT_POSEIDON, ps, Pol::single(psdn::FILTER_LOOKED_NORMAL))));
    }
    // 11. ctl_cpu_poseidon_tree_key (:415-429): CPU -> Poseidon(input, output[0..4] | filter_looked_treekey) (poseidon_stark.rs:151-157)
    {
        Cols c = S({COL_ADDR_STORAGE, COL_ADDR_STORAGE + 1, COL_ADDR_STORAGE + 2, COL_ADDR_STORAGE + 3, COL_S_OP0 + 4, COL_S_OP0 + 5, COL_S_OP0 + 6, COL_S_OP0 + 7});
        for (int i = 0; i < 4; ++i) c.push_back(Pol::linear({}, 0));  // Column::zero()
        for (int i = 0; i < 4; ++i) c.push_back(Pol::single(COL_S_DST + i));
        Cols ps;
        for (int i = 0; i < 16; ++i) ps.push_back(Pol::single(psdn::COL_POSEIDON_INPUT + i));
        v.push_back(full({Pol::twc(RT_CPU, c, Pol::single(COL_IS_STORAGE_EXT_LINE))}, Pol::twc(RT_POSEIDON, ps, Pol::single(psdn::FILTER_LOOKED_TREEKEY))));
    }
    // 12. ctl_cpu_storage_access (:372-386): CPU -> StorageAccess(idx, is_write, addr, path | is_layer_256 - filter_is_for_prog)
    //     (storage_access_stark.rs:33-49)
    {
        using namespace storage;
        Cols st = S({COL_ST_ACCESS_IDX, COL_ST_IS_WRITE});
        for (int i = 0; i < 4; ++i) st.push_back(Pol::single(COL_ST_ADDR + i));
        for (int i = 0; i < 4; ++i) st.push_back(Pol::single(COL_ST_PATH + i));
        v.push_back(full({Pol::twc(RT_CPU, S({COL_IDX_STORAGE, COL_S_SSTORE, COL_S_DST, COL_S_DST + 1, COL_S_DST + 2, COL_S_DST + 3, COL_S_OP1 + 4, COL_S_OP1 + 5, COL_S_OP1 + 6, COL_S_OP1 + 7}),
                                   Pol::single(COL_IS_STORAGE_EXT_LINE))},
                         Pol::twc(RT_STORAGE, st, Pol::linear({{COL_ST_IS_LAYER_256, 1}, {COL_ST_FILTER_IS_FOR_PROG, NEG_ONE}}, 0))));
    }
    // 13. ctl_storage_access_poseidon (:388-413): 4 StorageAccess lookers (storage_access_stark.rs:51-107) ->
    //     Poseidon(input, output[0..4], filter_leaf, filter_branch | filter_leaf + filter_branch) (poseidon_stark.rs:167-179)
    {
        using namespace storage;
        auto looker = [&](int first, int second, int hash, int filter) {
            Cols c;
            for (int i = 0; i < 4; ++i) c.push_back(Pol::single(first + i));
            for (int i = 0; i < 4; ++i) c.push_back(Pol::single(second + i));
            c.push_back(Pol::single(COL_ST_HASH_TYPE));
            for (int i = 0; i < 3; ++i) c.push_back(Pol::linear({}, 0));
            for (int i = 0; i < 4; ++i) c.push_back(Pol::single(hash + i));
            c.push_back(Pol::single(COL_ST_IS_LAYER_256));
            c.push_back(Pol::linear({{COL_ST_IS_LAYER_256, NEG_ONE}}, 1));
            return Pol::twc(RT_STORAGE, c, Pol::single(filter));
        };
        std::vector<Twc> l;
        l.push_back(looker(COL_ST_PATH, COL_ST_SIB, COL_ST_HASH, COL_ST_FILTER_IS_HASH_BIT_0));
        l.push_back(looker(COL_ST_PRE_PATH, COL_ST_SIB, COL_ST_PRE_HASH, COL_ST_FILTER_IS_HASH_BIT_0));
        l.push_back(looker(COL_ST_SIB, COL_ST_PATH, COL_ST_HASH, COL_ST_FILTER_IS_HASH_BIT_1));
        l.push_back(looker(COL_ST_SIB, COL_ST_PRE_PATH, COL_ST_PRE_HASH, COL_ST_FILTER_IS_HASH_BIT_1));
        Cols ps;
        for (int i = 0; i < 16; ++i) ps.push_back(Pol::single(psdn::COL_POSEIDON_INPUT + i));
        ps.push_back(Pol::single(psdn::FILTER_LOOKED_STORAGE_LEAF));
        ps.push_back(Pol::single(psdn::FILTER_LOOKED_STORAGE_BRANCH));
        v.push_back(full(l, Pol::twc(RT_POSEIDON, ps, sum({psdn::FILTER_LOOKED_STORAGE_LEAF, psdn::FILTER_LOOKED_STORAGE_BRANCH}))));
    }
    // 14. ctl_cpu_tape (:431-474): 13 CPU lookers -> Tape(tx, opcode, addr, value | filter_looked) (tape_stark.rs:25-38)
    {
        std::vector<Twc> l;
        l.push_back(Pol::twc(RT_CPU, S({COL_TX_IDX, COL_OPCODE, COL_S_OP0, COL_AUX1}), Pol::single(COL_FILTER_TAPE_LOOKING)));
        for (int i = 0; i < 4; ++i) {
            Cols c = S({COL_TX_IDX, COL_OPCODE});
            c.push_back(plus(COL_TP, (uint64_t)i));
            c.push_back(Pol::single(COL_S_OP0 + i));
            l.push_back(Pol::twc(RT_CPU, c, Pol::single(IS_SCCALL_EXT_LINE)));
        }
        for (int i = 0; i < 4; ++i) {
            Cols c = S({COL_TX_IDX, COL_OPCODE});
            c.push_back(plus(COL_TP, (uint64_t)(4 + i)));
            c.push_back(Pol::single(COL_ADDR_CODE + i));
            l.push_back(Pol::twc(RT_CPU, c, Pol::single(IS_SCCALL_EXT_LINE)));
        }
        for (int i = 0; i < 4; ++i) {
            Cols c = S({COL_TX_IDX, COL_OPCODE});
            c.push_back(plus(COL_TP, (uint64_t)(8 + i)));
            c.push_back(Pol::single(COL_ADDR_STORAGE + i));
            l.push_back(Pol::twc(RT_CPU, c, Pol::single(IS_SCCALL_EXT_LINE)));
        }
        v.push_back(full(l, Pol::twc(RT_TAPE, S({tape::COL_TAPE_TX_IDX, tape::COL_TAPE_OPCODE, tape::COL_TAPE_ADDR, tape::COL_TAPE_VALUE}), Pol::single(tape::COL_FILTER_LOOKED))));
    }
    // 15. ctl_cpu_sccall (:476-490): CPU -> SCCall (sccall_stark.rs:22-40)
    {
        using namespace sccall;
        Cols c = S({COL_TX_IDX, COL_ENV_IDX});
        for (int i = 0; i < 8; ++i) c.push_back(Pol::single(COL_S_OP0 + i));
        c.push_back(Pol::single(COL_CLK));
        c.push_back(Pol::single(COL_OP1_IMM));
        for (int i = 0; i < REGISTER_NUM; ++i) c.push_back(Pol::single(COL_REGS + i));
        c.push_back(plus(COL_ENV_IDX, 1));
        Cols sc = S({COL_SCCALL_TX_IDX, COL_SCCALL_CALLER_ENV_IDX});
        for (int i = 0; i < 4; ++i) sc.push_back(Pol::single(COL_SCCALL_CALLER_EXE_CTX + i));
        for (int i = 0; i < 4; ++i) sc.push_back(Pol::single(COL_SCCALL_CALLER_CODE_CTX + i));
        sc.push_back(Pol::single(COL_SCCALL_CLK_CALLER_CALL));
        sc.push_back(Pol::single(COL_SCCALL_CALLER_OP1_IMM));
        for (int i = 0; i < 10; ++i) sc.push_back(Pol::single(COL_SCCALL_CALLER_REG + i));
        sc.push_back(Pol::single(COL_SCCALL_CALLEE_ENV_IDX));
        v.push_back(full({Pol::twc(RT_CPU, c, Pol::single(IS_SCCALL_EXT_LINE))}, Pol::twc(RT_SCCALL, sc, Pol::linear({{COL_SCCALL_IS_PADDING, NEG_ONE}}, 1))));
    }
    // 16. ctl_cpu_sccall_end (:492-506): CPU -> SCCall (sccall_stark.rs:42-60)
    {
        using namespace sccall;
        Cols c = S({COL_TX_IDX, COL_ENV_IDX});
        for (int i = 0; i < 4; ++i) c.push_back(Pol::single(COL_ADDR_STORAGE + i));
        for (int i = 0; i < 4; ++i) c.push_back(Pol::single(COL_ADDR_CODE + i));
        c.push_back(Pol::single(COL_CLK));
        for (int i = 0; i < REGISTER_NUM; ++i) c.push_back(Pol::single(COL_REGS + i));
        c.push_back(Pol::single(COL_AUX0));
        c.push_back(Pol::single(COL_AUX1));
        Cols sc = S({COL_SCCALL_TX_IDX, COL_SCCALL_CALLER_ENV_IDX});
        for (int i = 0; i < 4; ++i) sc.push_back(Pol::single(COL_SCCALL_CALLER_EXE_CTX + i));
        for (int i = 0; i < 4; ++i) sc.push_back(Pol::single(COL_SCCALL_CALLER_CODE_CTX + i));
        sc.push_back(Pol::single(COL_SCCALL_CLK_CALLER_CALL));
        for (int i = 0; i < 10; ++i) sc.push_back(Pol::single(COL_SCCALL_CALLER_REG + i));
        sc.push_back(Pol::single(COL_SCCALL_CALLEE_ENV_IDX));
        sc.push_back(Pol::single(COL_SCCALL_CLK_CALLEE_END));
        v.push_back(full({Pol::twc(RT_CPU, c, Pol::single(COL_FILTER_SCCALL_END))}, Pol::twc(RT_SCCALL, sc, Pol::linear({{COL_SCCALL_IS_PADDING, NEG_ONE}}, 1))));
    }
    // 17. ctl_cpu_program (:508-528): CPU x2 -> Program(exec_code_addr, exec_pc, exec_inst | filter_exec) (program_stark.rs:24-30)
    {
        using namespace program;
        Cols inst = S({COL_ADDR_CODE, COL_ADDR_CODE + 1, COL_ADDR_CODE + 2, COL_ADDR_CODE + 3, COL_PC, COL_INST});
        Cols imm = S({COL_ADDR_CODE, COL_ADDR_CODE + 1, COL_ADDR_CODE + 2, COL_ADDR_CODE + 3});
        imm.push_back(plus(COL_PC, 1));
        imm.push_back(Pol::single(COL_IMM_VAL));
        v.push_back(full({Pol::twc(RT_CPU, inst, Pol::linear({{COL_IS_EXT_LINE, NEG_ONE}, {COL_IS_PADDING, NEG_ONE}}, 1)),
                          Pol::twc(RT_CPU, imm, Pol::single(COL_FILTER_LOOKING_PROG_IMM))},
                         Pol::twc(RT_PROGRAM, S({COL_PROG_EXEC_CODE_ADDR, COL_PROG_EXEC_CODE_ADDR + 1, COL_PROG_EXEC_CODE_ADDR + 2, COL_PROG_EXEC_CODE_ADDR + 3, COL_PROG_EXEC_PC, COL_PROG_EXEC_INST}),
                                  Pol::single(COL_PROG_FILTER_EXEC))));
    }
    // 18. ctl_prog_chunk_prog (:530-547): 8 ProgChunk lookers (prog_chunk_stark.rs:23-37) -> Program(code_addr, pc, inst | filter_prog_chunk)
    {
        using namespace prog_chunk;
        std::vector<Twc> l;
        for (int i = 0; i < 8; ++i) {
            Cols c = S({COL_PROG_CHUNK_CODE_ADDR, COL_PROG_CHUNK_CODE_ADDR + 1, COL_PROG_CHUNK_CODE_ADDR + 2, COL_PROG_CHUNK_CODE_ADDR + 3});
            c.push_back(plus(COL_PROG_CHUNK_START_PC, (uint64_t)i));
            c.push_back(Pol::single(COL_PROG_CHUNK_INST + i));
            l.push_back(Pol::twc(RT_PROG_CHUNK, c, Pol::single(COL_PROG_CHUNK_FILTER_LOOKING_PROG + i)));
        }
        using namespace program;
        v.push_back(full(l, Pol::twc(RT_PROGRAM, S({COL_PROG_CODE_ADDR, COL_PROG_CODE_ADDR + 1, COL_PROG_CODE_ADDR + 2, COL_PROG_CODE_ADDR + 3, COL_PROG_PC, COL_PROG_INST}),
                                     Pol::single(COL_PROG_FILTER_PROG_CHUNK))));
    }
    // 19. ctl_prog_chunk_storage (:549-563): ProgChunk(0, code_addr, hash[0..4] | is_result_line) (prog_chunk_stark.rs:52-62)
    //     -> StorageAccess(is_write, addr, path | filter_is_for_prog) (storage_access_stark.rs:24-31)
    {
        using namespace prog_chunk;
        using namespace storage;
        Cols c;
        c.push_back(Pol::linear({}, 0));
        for (int i = 0; i < 4; ++i) c.push_back(Pol::single(COL_PROG_CHUNK_CODE_ADDR + i));
        for (int i = 0; i < 4; ++i) c.push_back(Pol::single(COL_PROG_CHUNK_HASH + i));
        Cols st = S({COL_ST_IS_WRITE});
        for (int i = 0; i < 4; ++i) st.push_back(Pol::single(COL_ST_ADDR + i));
        for (int i = 0; i < 4; ++i) st.push_back(Pol::single(COL_ST_PATH + i));
        v.push_back(full({Pol::twc(RT_PROG_CHUNK, c, Pol::single(COL_PROG_CHUNK_IS_RESULT_LINE))}, Pol::twc(RT_STORAGE, st, Pol::single(COL_ST_FILTER_IS_FOR_PROG))));
    }
    return v;
}

}  // namespace air
}  // namespace ola
