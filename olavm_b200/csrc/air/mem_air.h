// AIR constraints of the Memory table (MemoryStark, 29 columns, constraint degree 8), single transcription shared
// by the device quotient kernel and the oracle (see the note at the top of cpu_air.h).
//   circuits/src/memory/memory_stark.rs:92-340 eval_packed_generic (constraints in source order)
//   circuits/src/memory/columns.rs:13-45 column indices;  ADDR_HEAP_PTR / INIT_VALUE_HEAP_PTR :79-80
// The reference's `is_next_addr_heap_ptr` (:290-298) is a data-dependent branch ("all lanes of the packed value equal
// the heap-pointer address"); with P::WIDTH == 1 it is 1 iff nv_addr == ADDR_HEAP_PTR.
#pragma once
#include "cpu_air.h"

namespace ola {
namespace air {

// exact zero test of a field value; specialised by each side
template <class T>
AIR_FN bool is_zero(const T& x);

namespace mem {
enum : int {
    COL_MEM_TX_IDX = 0, COL_MEM_ENV_IDX, COL_MEM_IS_RW, COL_MEM_ADDR, COL_MEM_CLK, COL_MEM_OP, COL_MEM_S_MLOAD, COL_MEM_S_MSTORE, COL_MEM_S_CALL,
    COL_MEM_S_RET, COL_MEM_S_TLOAD, COL_MEM_S_TSTORE, COL_MEM_S_SCCALL, COL_MEM_S_POSEIDON, COL_MEM_S_SSTORE, COL_MEM_S_SLOAD, COL_MEM_S_PROPHET,
    COL_MEM_IS_WRITE, COL_MEM_VALUE, COL_MEM_DIFF_ADDR, COL_MEM_DIFF_ADDR_INV, COL_MEM_DIFF_CLK, COL_MEM_DIFF_ADDR_COND, COL_MEM_RW_ADDR_UNCHANGED,
    COL_MEM_REGION_PROPHET, COL_MEM_REGION_HEAP, COL_MEM_RC_VALUE, COL_MEM_FILTER_LOOKING_RC, COL_MEM_FILTER_LOOKING_RC_COND, NUM_MEM_COLS
};
static constexpr uint64_t ADDR_HEAP_PTR = 18446744060824649731ULL;
static constexpr uint64_t INIT_VALUE_HEAP_PTR = ADDR_HEAP_PTR + 1;

template <class T, class R, class C>
AIR_FN void eval(const R& lv, const R& nv, C& yc) {
    using namespace cpu;  // opcode masks
    const T one = kc<T>(1);
    yc.constraint_transition((nv[COL_MEM_TX_IDX] - lv[COL_MEM_TX_IDX]) * (one - nv[COL_MEM_TX_IDX] + lv[COL_MEM_TX_IDX]));
    yc.constraint_transition((one - nv[COL_MEM_TX_IDX] + lv[COL_MEM_TX_IDX]) * (nv[COL_MEM_ENV_IDX] - lv[COL_MEM_ENV_IDX]) *
                             (one - nv[COL_MEM_ENV_IDX] + lv[COL_MEM_ENV_IDX]));
    const T p = kc<T>(0);
    const T span = kc<T>(0xFFFFFFFFULL);
    const T addr_heap_ptr = kc<T>(ADDR_HEAP_PTR);
    const T is_rw = lv[COL_MEM_IS_RW];
    const T region_prophet = lv[COL_MEM_REGION_PROPHET], nv_region_prophet = nv[COL_MEM_REGION_PROPHET];
    const T region_heap = lv[COL_MEM_REGION_HEAP], nv_region_heap = nv[COL_MEM_REGION_HEAP];
    const T region_stack = one - lv[COL_MEM_REGION_HEAP] - lv[COL_MEM_REGION_PROPHET];
    const T nv_region_stack = one - nv[COL_MEM_REGION_HEAP] - nv[COL_MEM_REGION_PROPHET];
    const T is_write = lv[COL_MEM_IS_WRITE], nv_is_write = nv[COL_MEM_IS_WRITE];
    const T addr = lv[COL_MEM_ADDR], nv_addr = nv[COL_MEM_ADDR];
    const T nv_diff_addr_inv = nv[COL_MEM_DIFF_ADDR_INV];
    const T diff_addr = lv[COL_MEM_DIFF_ADDR], nv_diff_addr = nv[COL_MEM_DIFF_ADDR];
    const T rw_addr_unchanged = lv[COL_MEM_RW_ADDR_UNCHANGED], nv_rw_addr_unchanged = nv[COL_MEM_RW_ADDR_UNCHANGED];
    const T diff_addr_cond = lv[COL_MEM_DIFF_ADDR_COND];
    const T value = lv[COL_MEM_VALUE], nv_value = nv[COL_MEM_VALUE];
    const T diff_clk = lv[COL_MEM_DIFF_CLK];
    const T rc_value = lv[COL_MEM_RC_VALUE];
    const T filter_looking_rc = lv[COL_MEM_FILTER_LOOKING_RC];
    const T lv_filter_looking_rc_cond = lv[COL_MEM_FILTER_LOOKING_RC_COND];

    const int sel[11] = {COL_MEM_S_MLOAD, COL_MEM_S_MSTORE, COL_MEM_S_CALL, COL_MEM_S_RET, COL_MEM_S_TLOAD, COL_MEM_S_TSTORE, COL_MEM_S_SCCALL,
                         COL_MEM_S_POSEIDON, COL_MEM_S_SSTORE, COL_MEM_S_SLOAD, COL_MEM_S_PROPHET};
    const uint64_t opc[11] = {OP_MLOAD, OP_MSTORE, OP_CALL, OP_RET, OP_TLOAD, OP_TSTORE, OP_SCCALL, OP_POSEIDON, OP_SSTORE, OP_SLOAD, 0};
    for (int i = 0; i < 11; ++i) yc.constraint((lv[COL_MEM_OP] - kc<T>(opc[i])) * lv[sel[i]]);
    for (int i = 0; i < 11; ++i) yc.constraint((one - lv[sel[i]]) * lv[sel[i]]);
    {
        T s = one;
        for (int i = 0; i < 11; ++i) s = s - lv[sel[i]];
        yc.constraint(s);
    }
    // is_rw region
    yc.constraint(is_rw * (one - is_rw));
    yc.constraint(lv[COL_MEM_IS_RW] * lv[COL_MEM_S_PROPHET]);
    yc.constraint((one - lv[COL_MEM_IS_RW]) * (one - lv[COL_MEM_S_PROPHET] - lv[COL_MEM_S_MLOAD]));
    // is_write
    yc.constraint(lv[COL_MEM_IS_WRITE] *
                  (one - lv[COL_MEM_S_MSTORE] - lv[COL_MEM_S_CALL] - lv[COL_MEM_S_TLOAD] - lv[COL_MEM_S_POSEIDON] - lv[COL_MEM_S_SLOAD] - lv[COL_MEM_S_PROPHET]));
    yc.constraint((one - lv[COL_MEM_IS_WRITE]) * (one - lv[COL_MEM_S_MLOAD] - lv[COL_MEM_S_CALL] - lv[COL_MEM_S_RET] - lv[COL_MEM_S_TSTORE] -
                                                 lv[COL_MEM_S_SCCALL] - lv[COL_MEM_S_POSEIDON] - lv[COL_MEM_S_SSTORE] - lv[COL_MEM_S_SLOAD]));
    // regions
    yc.constraint(one - region_stack - region_heap - region_prophet);
    yc.constraint(region_stack * (one - region_stack));
    yc.constraint(region_heap * (one - region_heap));
    yc.constraint(region_prophet * (one - region_prophet));
    yc.constraint(region_prophet * (p - addr - diff_addr_cond));
    yc.constraint(region_heap * (p - span - addr - diff_addr_cond));
    const T same_tx = one - nv[COL_MEM_TX_IDX] + lv[COL_MEM_TX_IDX];
    const T same_env = one - nv[COL_MEM_ENV_IDX] + lv[COL_MEM_ENV_IDX];
    yc.constraint_transition(same_tx * same_env * (nv_region_heap - region_heap - one) * (nv_addr - addr - nv_diff_addr));
    yc.constraint_transition(same_tx * same_env * region_stack * nv_region_stack * (one - nv_rw_addr_unchanged - nv_diff_addr * nv_diff_addr_inv));
    yc.constraint_transition(same_tx * same_env * region_heap * nv_region_heap * (one - nv_rw_addr_unchanged - nv_diff_addr * nv_diff_addr_inv));
    // write once
    yc.constraint(region_prophet * nv_region_prophet * (nv_addr - addr) * (nv_addr - addr - one));
    yc.constraint(region_prophet * nv_region_prophet * (nv_addr - addr - one) * nv_is_write);
    // read/write
    yc.constraint_first_row(is_rw * (one - is_write) * (addr - addr_heap_ptr));
    yc.constraint((nv[COL_MEM_TX_IDX] - lv[COL_MEM_TX_IDX]) * (nv[COL_MEM_ENV_IDX] - lv[COL_MEM_ENV_IDX]) * nv[COL_MEM_IS_RW] * (one - nv_is_write) *
                  (nv_addr - addr_heap_ptr));
    yc.constraint((nv_addr - addr) * (one - nv_is_write) * (nv_addr - addr_heap_ptr));
    yc.constraint((one - nv_is_write) * (nv_value - value) * (nv_addr - addr_heap_ptr));
    const T is_next_addr_heap_ptr = is_zero<T>(nv_addr - addr_heap_ptr) ? one : kc<T>(0);
    yc.constraint(is_next_addr_heap_ptr * (nv_addr - addr_heap_ptr));
    yc.constraint((addr - addr_heap_ptr) * is_next_addr_heap_ptr * (one - nv_is_write) * (nv_value - kc<T>(INIT_VALUE_HEAP_PTR)));
    // rc_value
    yc.constraint_transition(same_tx * same_env * is_rw * (nv_region_heap - region_heap - one) * (rc_value - rw_addr_unchanged * diff_clk) *
                             (rc_value - (one - rw_addr_unchanged) * diff_addr));
    yc.constraint_transition(same_tx * same_env * is_rw * rc_value * (nv_region_heap - region_heap - one) * (one - filter_looking_rc));
    yc.constraint((one - lv_filter_looking_rc_cond) * region_heap);
    yc.constraint((one - lv_filter_looking_rc_cond) * region_prophet * (one - is_write));
}
}  // namespace mem
}  // namespace air
}  // namespace ola
