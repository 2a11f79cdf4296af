"""The data-parallel derivation of `permuted_cols` (olavm_b200/csrc/lookup.cu) checked on the CPU, step for step, against the
oracle's statement-by-statement restatement of the reference's serial merge walk (circuits/src/stark/lookup.rs:68-131).

The GPU kernels implement exactly the steps below (sorts, binary searches, two prefix sums for the event order, a running
+1 / -1 sum for the stack depth, a sort by (level, position) for the bracket matching, two compactions); this file is the
executable statement of WHY those steps reproduce the walk, independent of CUDA.  tests/test_generation.py compares the
kernels themselves with the oracle on the GPU."""
import numpy as np
import pytest

P = 0xFFFFFFFF00000001


def parallel_permuted_cols(inputs, table):
    n = len(inputs)
    S = np.sort(np.asarray(inputs, dtype=np.uint64) % np.uint64(P))
    T = np.sort(np.asarray(table, dtype=np.uint64) % np.uint64(P))
    idx = np.arange(n)
    # step 2: pairing by rank among equal values
    lbS_S, lbT_S, ubT_S = np.searchsorted(S, S, "left"), np.searchsorted(T, S, "left"), np.searchsorted(T, S, "right")
    paired_in = (idx - lbS_S) < (ubT_S - lbT_S)
    PT = np.where(paired_in, S, 0).astype(np.uint64)
    pop = ~paired_in & (S < T[-1])       # an unpaired input below max(table) pops the list of skipped table values ...
    hole = ~paired_in & (S >= T[-1])     # ... at or above it the walk has ended: the row is filled at the very end
    lbT_T, lbS_T, ubS_T = np.searchsorted(T, T, "left"), np.searchsorted(S, T, "left"), np.searchsorted(S, T, "right")
    push = ~((idx - lbT_T) < (ubS_T - lbS_T))   # an unpaired table entry is pushed
    # step 3: walk order of the events from two prefix sums (pushes and pops never share a value)
    cum_pop = np.concatenate([[0], np.cumsum(pop)])
    cum_push = np.concatenate([[0], np.cumsum(push)])
    ne = int(cum_pop[-1] + cum_push[-1])
    delta = np.zeros(ne, dtype=np.int64)
    ref = np.zeros(ne, dtype=np.int64)
    pi = np.nonzero(pop)[0]
    e = cum_pop[pi] + cum_push[lbT_S[pi]]
    delta[e], ref[e] = -1, n + pi
    pj = np.nonzero(push)[0]
    e = cum_push[pj] + cum_pop[lbS_T[pj]]
    delta[e], ref[e] = 1, pj
    assert (delta != 0).all()            # the positions are a permutation of 0 .. ne-1
    # step 4: bracket matching by (level, position)
    depth_before = np.concatenate([[0], np.cumsum(delta)])[:-1]
    level = depth_before + (delta == 1)
    order = np.lexsort((np.arange(ne), level))
    left = np.zeros(n, dtype=bool)
    hole = hole.copy()
    for s in range(ne):
        ev = order[s]
        if delta[ev] == 1:
            popped = s + 1 < ne and level[order[s + 1]] == level[ev] and delta[order[s + 1]] != 1
            left[ref[ev]] = not popped
        else:
            i = ref[ev] - n
            if s > 0 and level[order[s - 1]] == level[ev] and delta[order[s - 1]] == 1:
                PT[i] = T[ref[order[s - 1]]]
            else:
                hole[i] = True
    # step 5: the unfilled rows take the values never popped, both in order
    lv, hi = T[left], np.nonzero(hole)[0]
    assert len(lv) == len(hi)
    PT[hi] = lv
    return S, PT


def _case(rng, kind, n):
    if kind == 0:
        return rng.integers(0, n, size=n, dtype=np.uint64), np.arange(n, dtype=np.uint64)
    if kind == 1:
        return rng.integers(0, n // 2 + 1, size=n, dtype=np.uint64), np.minimum(np.arange(n), n // 2).astype(np.uint64)
    if kind == 2:
        return rng.integers(0, 12, size=n, dtype=np.uint64), rng.integers(0, 8, size=n, dtype=np.uint64)
    if kind == 3:
        return rng.integers(0, P, size=n, dtype=np.uint64), rng.integers(0, P, size=n, dtype=np.uint64)
    if kind == 4:
        return rng.integers(0, 16, size=n, dtype=np.uint64), rng.integers(0, 4, size=n, dtype=np.uint64) + np.uint64(5)
    if kind == 5:
        t = rng.integers(0, 6, size=n, dtype=np.uint64)
        a = t.copy()
        rng.shuffle(a)
        return a, t
    return rng.integers(0, 20, size=n, dtype=np.uint64), rng.integers(0, 5, size=n, dtype=np.uint64) * np.uint64(3)


@pytest.mark.parametrize("kind", range(7))
def test_parallel_derivation_equals_the_serial_walk(orc, kind):
    rng = np.random.default_rng(100 + kind)
    for trial in range(120):
        n = 1 << int(rng.integers(1, 8))
        a, t = _case(rng, kind, n)
        si, pt = parallel_permuted_cols(a, t)
        ri, rt = orc.permuted_cols(a, t)
        assert (si == ri).all() and (pt == rt).all(), (kind, trial, n)
