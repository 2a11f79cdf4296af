/* ORACLE (test infrastructure, NOT product code) -- the Memory table's AIR, restated from the Rust on its own:
 * circuits/src/memory/memory_stark.rs:92-340 (MemoryStark::eval_packed_generic), columns circuits/src/memory/columns.rs:4-39,
 * constants memory_stark.rs:81-82.  Constraints in source order.  `is_next_addr_heap_ptr` is the reference's value-level
 * branch (all lanes of the packed difference zero): with one row per evaluation it is "next address == ADDR_HEAP_PTR". */
#ifndef ORC_AIR_MEMORY_HPP
#define ORC_AIR_MEMORY_HPP
#include "stark.hpp"

namespace orc {
inline bool value_is_zero(F x) { return gl_canon(x) == 0; }
inline bool value_is_zero(E x) { return gl_canon(x.c0) == 0 && gl_canon(x.c1) == 0; }

namespace memory_air {
enum {
    TX_IDX = 0, ENV_IDX, IS_RW, ADDR, CLK, OP, S_MLOAD, S_MSTORE, S_CALL, S_RET, S_TLOAD, S_TSTORE, S_SCCALL, S_POSEIDON, S_SSTORE, S_SLOAD,
    S_PROPHET, IS_WRITE, VALUE, DIFF_ADDR, DIFF_ADDR_INV, DIFF_CLK, DIFF_ADDR_COND, RW_ADDR_UNCHANGED, REGION_PROPHET, REGION_HEAP, RC_VALUE,
    FILTER_LOOKING_RC, FILTER_LOOKING_RC_COND, NUM_COLS
};
static_assert(NUM_COLS == 29, "memory/columns.rs layout");
static const uint64_t ADDR_HEAP_PTR = 18446744060824649731ull;
static const uint64_t INIT_VALUE_HEAP_PTR = ADDR_HEAP_PTR + 1;

template <class O>
void eval(const P<O>* lv, const P<O>* nv, Consumer<O>& yc) {
    typedef P<O> T;
    const T ONE = T::one();
    const T same_tx = ONE - nv[TX_IDX] + lv[TX_IDX];   /* 1 - (next tx - local tx) */
    const T same_env = ONE - nv[ENV_IDX] + lv[ENV_IDX];
    yc.constraint_transition((nv[TX_IDX] - lv[TX_IDX]) * same_tx);
    yc.constraint_transition(same_tx * (nv[ENV_IDX] - lv[ENV_IDX]) * same_env);

    const T p = T::zero();
    const T span = T::c(((uint64_t)1 << 32) - 1);
    const T addr_heap_ptr = T::c(ADDR_HEAP_PTR);
    const T is_rw = lv[IS_RW];
    const T region_prophet = lv[REGION_PROPHET], nv_region_prophet = nv[REGION_PROPHET];
    const T region_heap = lv[REGION_HEAP], nv_region_heap = nv[REGION_HEAP];
    const T region_stack = ONE - lv[REGION_HEAP] - lv[REGION_PROPHET];
    const T nv_region_stack = ONE - nv[REGION_HEAP] - nv[REGION_PROPHET];
    const T is_write = lv[IS_WRITE], nv_is_write = nv[IS_WRITE];
    const T addr = lv[ADDR], nv_addr = nv[ADDR];
    const T nv_diff_addr_inv = nv[DIFF_ADDR_INV];
    const T diff_addr = lv[DIFF_ADDR], nv_diff_addr = nv[DIFF_ADDR];
    const T rw_addr_unchanged = lv[RW_ADDR_UNCHANGED], nv_rw_addr_unchanged = nv[RW_ADDR_UNCHANGED];
    const T diff_addr_cond = lv[DIFF_ADDR_COND];
    const T value = lv[VALUE], nv_value = nv[VALUE];
    const T diff_clk = lv[DIFF_CLK];
    const T rc_value = lv[RC_VALUE];
    const T filter_looking_rc = lv[FILTER_LOOKING_RC];
    const T filter_looking_rc_cond = lv[FILTER_LOOKING_RC_COND];

    /* opcode <-> selector: mload 22, mstore 21, call 24, ret 23, tload 9, tstore 8, sccall 7, poseidon 12, sstore 10, sload 11,
     * prophet writes carry opcode 0 */
    const int sel[11] = {S_MLOAD, S_MSTORE, S_CALL, S_RET, S_TLOAD, S_TSTORE, S_SCCALL, S_POSEIDON, S_SSTORE, S_SLOAD, S_PROPHET};
    const int bit[11] = {22, 21, 24, 23, 9, 8, 7, 12, 10, 11, -1};
    for (int i = 0; i < 11; i++) yc.constraint((lv[OP] - (bit[i] < 0 ? T::zero() : T::c((uint64_t)1 << bit[i]))) * lv[sel[i]]);
    for (int i = 0; i < 11; i++) yc.constraint((ONE - lv[sel[i]]) * lv[sel[i]]);
    {
        T rest = ONE;
        for (int i = 0; i < 11; i++) rest = rest - lv[sel[i]];
        yc.constraint(rest);
    }
    yc.constraint(is_rw * (ONE - is_rw));
    yc.constraint(lv[IS_RW] * lv[S_PROPHET]);
    yc.constraint((ONE - lv[IS_RW]) * (ONE - lv[S_PROPHET] - lv[S_MLOAD]));
    yc.constraint(lv[IS_WRITE] * (ONE - lv[S_MSTORE] - lv[S_CALL] - lv[S_TLOAD] - lv[S_POSEIDON] - lv[S_SLOAD] - lv[S_PROPHET]));
    yc.constraint((ONE - lv[IS_WRITE]) *
                  (ONE - lv[S_MLOAD] - lv[S_CALL] - lv[S_RET] - lv[S_TSTORE] - lv[S_SCCALL] - lv[S_POSEIDON] - lv[S_SSTORE] - lv[S_SLOAD]));

    yc.constraint(ONE - region_stack - region_heap - region_prophet);
    yc.constraint(region_stack * (ONE - region_stack));
    yc.constraint(region_heap * (ONE - region_heap));
    yc.constraint(region_prophet * (ONE - region_prophet));
    yc.constraint(region_prophet * (p - addr - diff_addr_cond));
    yc.constraint(region_heap * (p - span - addr - diff_addr_cond));

    yc.constraint_transition(same_tx * same_env * (nv_region_heap - region_heap - ONE) * (nv_addr - addr - nv_diff_addr));
    yc.constraint_transition(same_tx * same_env * region_stack * nv_region_stack * (ONE - nv_rw_addr_unchanged - nv_diff_addr * nv_diff_addr_inv));
    yc.constraint_transition(same_tx * same_env * region_heap * nv_region_heap * (ONE - nv_rw_addr_unchanged - nv_diff_addr * nv_diff_addr_inv));

    yc.constraint(region_prophet * nv_region_prophet * (nv_addr - addr) * (nv_addr - addr - ONE));
    yc.constraint(region_prophet * nv_region_prophet * (nv_addr - addr - ONE) * nv_is_write);

    yc.constraint_first_row(is_rw * (ONE - is_write) * (addr - addr_heap_ptr));
    yc.constraint((nv[TX_IDX] - lv[TX_IDX]) * (nv[ENV_IDX] - lv[ENV_IDX]) * nv[IS_RW] * (ONE - nv_is_write) * (nv_addr - addr_heap_ptr));
    yc.constraint((nv_addr - addr) * (ONE - nv_is_write) * (nv_addr - addr_heap_ptr));
    yc.constraint((ONE - nv_is_write) * (nv_value - value) * (nv_addr - addr_heap_ptr));

    const T is_next_addr_heap_ptr = value_is_zero((nv_addr - T::c(ADDR_HEAP_PTR)).v) ? ONE : T::zero();
    yc.constraint(is_next_addr_heap_ptr * (nv_addr - T::c(ADDR_HEAP_PTR)));
    yc.constraint((addr - T::c(ADDR_HEAP_PTR)) * is_next_addr_heap_ptr * (ONE - nv_is_write) * (nv_value - T::c(INIT_VALUE_HEAP_PTR)));

    yc.constraint_transition(same_tx * same_env * is_rw * (nv_region_heap - region_heap - ONE) * (rc_value - rw_addr_unchanged * diff_clk) *
                             (rc_value - (ONE - rw_addr_unchanged) * diff_addr));
    yc.constraint_transition(same_tx * same_env * is_rw * rc_value * (nv_region_heap - region_heap - ONE) * (ONE - filter_looking_rc));

    yc.constraint((ONE - filter_looking_rc_cond) * region_heap);
    yc.constraint((ONE - filter_looking_rc_cond) * region_prophet * (ONE - is_write));
}
}  // namespace memory_air
}  // namespace orc
#endif
