"""GPU parity tests (-m gpu) of the STARK prover: proof BYTES from the CUDA path (through the C ABI) must equal the
oracle's for the same traces, and the oracle's restated verify_proof must accept them."""
import os

import numpy as np
import pytest

import olavm_b200
import tracegen

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001
CMP, RC = 3, 4


def _valid_cmp_rc(seed, log_cmp):
    rng = np.random.default_rng(seed)
    k = (1 << log_cmp) - 5
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(k, 2))] + [(5, 5), (0, 9)]
    return tracegen.cmp_trace(pairs, log_cmp), tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])


@pytest.mark.parametrize("seed,log_cmp", [(5, 6), (6, 4), (7, 9)])
def test_cmp_rangecheck_proof_bytes_equal_oracle(ctx, orc, seed, log_cmp):
    cmp_t, rc_t = _valid_cmp_rc(seed, log_cmp)
    ref = orc.stark_prove([CMP, RC], [cmp_t, rc_t])
    got = olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t])
    assert len(got) == len(ref)
    assert got == ref
    ok, msg = orc.stark_verify([CMP, RC], got)
    assert ok, msg


def test_single_table_system_has_only_partial_ctls(ctx, orc):
    # one table alone: every CTL it takes part in loses its other side ("partial"); its Z columns are still proven and
    # the bytes still match the oracle
    cmp_t, _ = _valid_cmp_rc(1, 5)
    assert olavm_b200.prove_with_traces(ctx, [CMP], [cmp_t]) == orc.stark_prove([CMP], [cmp_t])


def test_pipeline_parity_on_random_traces(ctx, orc):
    """Random (non-satisfying) columns with binary filters: prover must still agree byte for byte with the oracle
    (check_quotient_degree off = SURVEY section 7 option (i)); the verifier rejects such a proof."""
    rng = np.random.default_rng(99)
    cmp_t = rng.integers(0, P, size=(6, 1 << 7), dtype=np.uint64)
    cmp_t[5] = rng.integers(0, 2, size=1 << 7)
    rc_t = rng.integers(0, P, size=(12, 1 << 16), dtype=np.uint64)
    for c in range(4):
        rc_t[c] = rng.integers(0, 2, size=1 << 16)
    rc_t[0, 0] = np.uint64(P + 1)  # non-canonical representative of 1
    ref = orc.stark_prove([CMP, RC], [cmp_t, rc_t], check_degree=False)
    got = olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t], check_quotient_degree=False)
    assert got == ref
    ok, _ = orc.stark_verify([CMP, RC], got)
    assert not ok


def test_non_binary_filter_error(ctx):
    cmp_t, rc_t = _valid_cmp_rc(3, 5)
    cmp_t[5, 1] = 2
    with pytest.raises(olavm_b200.OlaError, match="Non-binary filter"):
        olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t])


CPU = 0


def test_cpu_table_valid_padding_trace(ctx, orc):
    cpu_t = tracegen.cpu_padding_trace(5)
    cmp_t = tracegen.cmp_trace([], 4)
    rc_t = tracegen.rangecheck_trace([])
    ref = orc.stark_prove([CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    got = olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    assert got == ref
    ok, msg = orc.stark_verify([CPU, CMP, RC], got)
    assert ok, msg
    bad = cpu_t.copy()
    bad[74, 3] = 0  # s_end must be 1 on padding rows
    with pytest.raises(olavm_b200.OlaError, match="Quotient has failed") as e:
        olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [bad, cmp_t, rc_t])
    assert e.value.code == -5  # OLA_ERR_QUOTIENT_DEGREE


def test_cpu_table_real_trace(ctx, orc):
    """A real CPU trace (tests/tracegen.py::cpu_vm_trace: rows produced the way the reference executor and
    generate_cpu_trace produce them, 10 opcodes, a taken and a fall-through branch, padding): the GPU prover's quotient
    passes the degree check, the proof bytes equal the oracle's and the restated verifier accepts; one wrong register
    is rejected with OLA_ERR_QUOTIENT_DEGREE."""
    cpu_t, steps = tracegen.cpu_vm_trace(tracegen.fib_program(10), 7)
    cmp_t = tracegen.cmp_trace([], 4)
    rc_t = tracegen.rangecheck_trace([])
    got = olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    assert got == orc.stark_prove([CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    ok, msg = orc.stark_verify([CPU, CMP, RC], got)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof([CPU, CMP, RC], got)
    assert ok, msg
    bad = cpu_t.copy()
    i_add = next(i for i, s in enumerate(steps) if s["op"] == "add")
    bad[16 + 3, i_add + 1] = (int(bad[16 + 3, i_add + 1]) + 1) % tracegen.P
    with pytest.raises(olavm_b200.OlaError, match="Quotient has failed") as e:
        olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [bad, cmp_t, rc_t])
    assert e.value.code == -5
    # a longer run of the same program at a larger table size (2^12 rows, 2053 executed steps)
    big, big_steps = tracegen.cpu_vm_trace(tracegen.fib_program(340), 12)
    assert len(big_steps) > 2000
    got = olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [big, cmp_t, rc_t])
    ok, msg = orc.stark_verify([CPU, CMP, RC], got)
    assert ok, msg


def test_cpu_cmp_rangecheck_real_program(ctx, orc):
    """calls_program (mstore / mload / call / ret / gte / range + arithmetic) with the Cmp and RangeCheck rows the executor
    would have inserted: real cpu->cmp, cmp->rangecheck and cpu->rangecheck lookups; GPU proof bytes equal the oracle's,
    both verifiers accept."""
    cpu_t, steps, cmp_pairs, rc_cmp, rc_cpu = tracegen.cpu_vm_trace(tracegen.calls_program(12), 9, want_side_tables=True)
    cmp_t = tracegen.cmp_trace(cmp_pairs, 6)
    rc_t = tracegen.rangecheck_trace(rc_cmp, cpu_vals=rc_cpu)
    got = olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    assert got == orc.stark_prove([CPU, CMP, RC], [cpu_t, cmp_t, rc_t])
    ok, msg = orc.stark_verify([CPU, CMP, RC], got)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof([CPU, CMP, RC], got)
    assert ok, msg
    # a Cmp row withheld: the proof is produced (each table is consistent) and both verifiers reject the lookup
    short = olavm_b200.prove_with_traces(ctx, [CPU, CMP, RC], [cpu_t, tracegen.cmp_trace(cmp_pairs[:-1], 6),
                                                               tracegen.rangecheck_trace(rc_cmp[:-1], cpu_vals=rc_cpu)])
    assert not orc.stark_verify([CPU, CMP, RC], short)[0]
    assert not olavm_b200.verify_subsystem_proof([CPU, CMP, RC], short)[0]


def test_cpu_memory_cmp_rangecheck_real_program(ctx, orc):
    """[Cpu, Memory, Cmp, RangeCheck] of a real program (tests/tracegen.py: VM + gen_memory_table restatement): degree check
    on, GPU proof bytes equal the oracle's, both verifiers accept; a read that returns another value than was written is
    rejected by the Memory AIR on the GPU as well."""
    cpu_t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log = tracegen.cpu_vm_trace(tracegen.calls_program(12), 9, want_side_tables="memory")
    mem_t, rc_sort = tracegen.memory_trace_from_log(mem_log, 7)
    cmp_t = tracegen.cmp_trace(cmp_pairs, 6)
    rc_t = tracegen.rangecheck_trace(rc_cmp, cpu_vals=rc_cpu, mem_sort_vals=rc_sort)
    ids = [CPU, 1, CMP, RC]
    got = olavm_b200.prove_with_traces(ctx, ids, [cpu_t, mem_t, cmp_t, rc_t])
    assert got == orc.stark_prove(ids, [cpu_t, mem_t, cmp_t, rc_t])
    ok, msg = orc.stark_verify(ids, got)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg
    m = mem_t.copy()
    k = next(i for i in range(1, m.shape[1]) if m[23, i] == 1 and m[17, i] == 0)
    m[18, k] = (int(m[18, k]) + 1) % tracegen.P
    with pytest.raises(olavm_b200.OlaError, match="Quotient has failed") as e:
        olavm_b200.prove_with_traces(ctx, ids, [cpu_t, m, cmp_t, rc_t])
    assert e.value.code == -5
    # a longer run: 2^13 CPU rows, 2^11 memory rows
    cpu_t, steps, cmp_pairs, rc_cmp, rc_cpu, mem_log = tracegen.cpu_vm_trace(tracegen.calls_program(300, linear=True), 13, want_side_tables="memory")
    mem_t, rc_sort = tracegen.memory_trace_from_log(mem_log, 11)
    big = olavm_b200.prove_with_traces(ctx, ids, [cpu_t, mem_t, tracegen.cmp_trace(cmp_pairs, 10),
                                                   tracegen.rangecheck_trace(rc_cmp, cpu_vals=rc_cpu, mem_sort_vals=rc_sort)])
    ok, msg = orc.stark_verify(ids, big)
    assert ok, msg


def test_real_program_run_systems(ctx, orc):
    """tracegen.real_program_system ([Cpu, Memory, Cmp, RangeCheck, Poseidon, StorageAccess, Program, ProgChunk] of one
    program run, thirteen lookups with real data): GPU proof bytes equal the oracle's, both verifiers accept; a longer run
    (2^13 CPU rows) verifies as well."""
    ids, traces, cc = tracegen.real_program_system(orc, np.random.default_rng(5))
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == orc.stark_prove(ids, traces, compress_challenges=cc)
    ok, msg = orc.stark_verify(ids, got)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg
    ids, traces, cc = tracegen.real_program_system(orc, np.random.default_rng(6), n_iter=300, linear=True, cpu_log=13, mem_log_n=11,
                                                   cmp_log=10, prog_log=13)
    big = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    ok, msg = orc.stark_verify(ids, big)
    assert ok, msg
    # eleven tables: and / or / xor rows with the Bitwise table, poseidon calls with PoseidonChunk, tstore / tload with Tape
    ids, traces, cc = tracegen.real_program_system(orc, np.random.default_rng(7), bitwise=True, poseidon=True, tape=True, mem_log_n=8)
    assert len(ids) == 11
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == orc.stark_prove(ids, traces, compress_challenges=cc)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg


@pytest.mark.parametrize("name", ["fibo_recursive", "call", "tape", "bitwise", "comparison", "range_check", "memory", "mem_gep", "context_fetch"])
def test_reference_programs_run_and_prove(ctx, orc, name):
    """The reference's own assembly test programs (tests/golden/ola_programs.json) run through the restated VM: the GPU
    prover's quotients pass the degree check, the proof bytes equal the oracle's and the product verifier accepts."""
    import json

    text = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ola_programs.json")))["programs"][name]
    ids, traces, cc, steps = tracegen.run_system(orc, np.random.default_rng(3), tracegen.parse_ola_asm(text),
                                                 init_tape=tracegen.CONTEXT_TAPE if name == "context_fetch" else ())
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == orc.stark_prove(ids, traces, compress_challenges=cc)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg


@pytest.mark.parametrize("log_n", [4, 8, 11])
def test_cpu_table_pipeline_parity(ctx, orc, log_n):
    """The 94-column CPU table with its 39 CTL instances (78 Z columns, 12 quotient chunks) on random columns with
    binary filters, CPU table alone (every CTL partial): proof bytes equal the oracle's."""
    rng = np.random.default_rng(1000 + log_n)
    cpu_t = tracegen.cpu_random_trace(rng, log_n)
    ref = orc.stark_prove([CPU], [cpu_t], check_degree=False)
    got = olavm_b200.prove_with_traces(ctx, [CPU], [cpu_t], check_quotient_degree=False)
    assert got == ref


MEM = 1


def test_cpu_memory_cmp_rangecheck_pipeline_parity(ctx, orc):
    """Four tables, 7 registered CTLs among them (cpu-memory with its 16 CPU lookers, memory-rc x2, cmp-cpu, cmp-rc,
    rc-cpu, ...): random columns with binary filters, proof bytes equal the oracle's."""
    rng = np.random.default_rng(4242)
    traces = [tracegen.cpu_random_trace(rng, 7), tracegen.memory_random_trace(rng, 6), tracegen.cmp_random_trace(rng, 5),
              tracegen.rangecheck_random_trace(rng)]
    ids = [CPU, MEM, CMP, RC]
    ref = orc.stark_prove(ids, traces, check_degree=False)
    got = olavm_b200.prove_with_traces(ctx, ids, traces, check_quotient_degree=False)
    assert got == ref


def test_memory_table_alone(ctx, orc):
    rng = np.random.default_rng(77)
    t = tracegen.memory_random_trace(rng, 9)
    assert olavm_b200.prove_with_traces(ctx, [MEM], [t], check_quotient_degree=False) == orc.stark_prove([MEM], [t], check_degree=False)


# ---------------------------------------------------------------------------------------------------------------------
# The other eight tables (Bitwise, Poseidon, PoseidonChunk, StorageAccess, Tape, SCCall, Program, ProgChunk)
# ---------------------------------------------------------------------------------------------------------------------
from test_oracle_stark import SINGLE, _valid_single  # noqa: E402  (valid traces shared with the CPU suite)

RANDOM = dict(bitwise=(2, tracegen.bitwise_random_trace), poseidon=(5, tracegen.poseidon_random_trace),
              poseidon_chunk=(6, tracegen.poseidon_chunk_random_trace), storage=(7, tracegen.storage_random_trace),
              tape=(8, tracegen.tape_random_trace), sccall=(9, tracegen.sccall_random_trace), program=(10, tracegen.program_random_trace),
              prog_chunk=(11, tracegen.prog_chunk_random_trace))


@pytest.mark.parametrize("name", SINGLE)
def test_valid_trace_of_each_remaining_table(ctx, orc, name):
    ids, traces, cc = _valid_single(orc, name)
    ref = orc.stark_prove(ids, traces, True, compress_challenges=cc)
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == ref
    ok, msg = orc.stark_verify(ids, got)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)  # product prover -> product verifier
    assert ok, msg


@pytest.mark.parametrize("name", sorted(RANDOM))
@pytest.mark.parametrize("log_n", [3, 7])
def test_pipeline_parity_of_each_remaining_table(ctx, orc, name, log_n):
    tid, gen = RANDOM[name]
    rng = np.random.default_rng(500 + 16 * tid + log_n)
    t = gen(rng, log_n)
    cc = [int(rng.integers(0, P, dtype=np.uint64))]
    ref = orc.stark_prove([tid], [t], check_degree=False, compress_challenges=cc)
    got = olavm_b200.prove_with_traces(ctx, [tid], [t], check_quotient_degree=False, compress_challenges=cc)
    assert got == ref


def test_sccall_degree_quirk_same_bytes(ctx, orc):
    rng = np.random.default_rng(8)
    t = tracegen.sccall_valid_trace(rng, 4, used=5)
    assert olavm_b200.prove_with_traces(ctx, [9], [t]) == orc.stark_prove([9], [t], True)


def test_five_table_hash_system(ctx, orc):
    rng = np.random.default_rng(3)
    ids, traces, cc = tracegen.hash_system_valid(orc, rng)
    ref = orc.stark_prove(ids, traces, True, compress_challenges=cc)
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == ref
    ok, msg = orc.stark_verify(ids, got)
    assert ok, msg
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg
    # a broken Poseidon round witness is caught by the device-side degree check like the reference's panic
    bad = [t.copy() for t in traces]
    bad[0][70, 1] = (int(bad[0][70, 1]) + 1) % P
    with pytest.raises(olavm_b200.OlaError, match="Quotient has failed"):
        olavm_b200.prove_with_traces(ctx, ids, bad, compress_challenges=cc)


def test_all_twelve_tables_pipeline_parity(ctx, orc):
    """prove_with_traces over the full Table enum with all 19 cross-table lookups (ola_stark.rs:121-143) on random
    columns with binary filters: proof bytes equal the oracle's."""
    rng = np.random.default_rng(2024)
    logs = {0: 6, 1: 5, 2: 4, 3: 4, 5: 3, 6: 5, 7: 4, 8: 3, 9: 2, 10: 5, 11: 4}
    traces = []
    for tid in range(12):
        if tid == 0:
            traces.append(tracegen.cpu_random_trace(rng, logs[0]))
        elif tid == 1:
            traces.append(tracegen.memory_random_trace(rng, logs[1]))
        elif tid == 3:
            traces.append(tracegen.cmp_random_trace(rng, logs[3]))
        elif tid == 4:
            traces.append(tracegen.rangecheck_random_trace(rng))
        else:
            name = [k for k, v in RANDOM.items() if v[0] == tid][0]
            traces.append(RANDOM[name][1](rng, logs[tid]))
    ids = list(range(12))
    cc = [int(x) for x in rng.integers(0, P, size=12, dtype=np.uint64)]
    ref = orc.stark_prove(ids, traces, check_degree=False, compress_challenges=cc)
    got = olavm_b200.prove_with_traces(ctx, ids, traces, check_quotient_degree=False, compress_challenges=cc)
    assert got == ref
    # the wire format's trailing compress_challenges carry only the Bitwise and Program entries (prover.rs:307-320)
    tail = np.frombuffer(got[-12 * 8:], dtype="<u8")
    assert [int(x) for x in tail] == [cc[i] if i in (2, 10) else 0 for i in range(12)]


# ---------------------------------------------------------------------------------------------------------------------
# Coset-sharded prover (ola_set_comm): `world` ranks as host threads on this one GPU, collectives through host staging.
# Every rank must return the single-GPU proof, byte for byte.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_prover_equals_single_gpu(ctx, orc, world):
    from olavm_b200 import dist as odist

    cmp_t, rc_t = _valid_cmp_rc(5, 6)
    single = olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t])
    proofs = odist.prove_sharded_local(0, world, [CMP, RC], [cmp_t, rc_t])
    assert all(p == single for p in proofs)
    ok, msg = orc.stark_verify([CMP, RC], proofs[-1])
    assert ok, msg


@pytest.mark.parametrize("hasher", [0, 1])
def test_sharded_prover_with_leaf_range_sharded_fri_layers(ctx, orc, hasher, monkeypatch):
    """Large FRI layers are committed by leaf range across the ranks (cap all-gathered, Merkle paths answered by the leaf's
    owner); the threshold is lowered so that this small system takes that path.  Same bytes as one GPU, both hashers."""
    from olavm_b200 import dist as odist

    monkeypatch.setenv("OLA_FRI_SHARD_MIN_LEAVES", "64")
    cmp_t, rc_t = _valid_cmp_rc(8, 9)   # RangeCheck 2^16 rows: first FRI layer 2^15 leaves, second 2^11, third 2^7
    ctx.hasher = hasher
    try:
        single = olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t])
    finally:
        ctx.hasher = 0
    for world in (2, 4, 8):
        proofs = odist.prove_sharded_local(0, world, [CMP, RC], [cmp_t, rc_t], hasher=hasher)
        assert all(p == single for p in proofs), world
    ok, msg = orc.stark_verify([CMP, RC], single, hasher_id=hasher)
    assert ok, msg


def test_sharded_prover_five_table_system_and_cpu_table(ctx, orc):
    from olavm_b200 import dist as odist

    rng = np.random.default_rng(3)
    ids, traces, cc = tracegen.hash_system_valid(orc, rng)  # degrees 3..7: quotient domains of 2, 4 and 8 cosets
    single = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    for world in (2, 8):
        proofs = odist.prove_sharded_local(0, world, ids, traces, compress_challenges=cc)
        assert all(p == single for p in proofs), world
    t = tracegen.cpu_random_trace(np.random.default_rng(12), 9)
    single = olavm_b200.prove_with_traces(ctx, [CPU], [t], check_quotient_degree=False)
    proofs = odist.prove_sharded_local(0, 4, [CPU], [t], check_quotient_degree=False)
    assert all(p == single for p in proofs)


def test_storage_opcodes_real_run(ctx, orc):
    """sstore / sload run through the VM (tests/tracegen.py::storage_program): Cpu storage ext lines, Memory, StorageAccess
    walking one consistent sparse Merkle tree, the tree-key / leaf / branch Poseidon rows: the GPU proof passes the degree
    check, equals the oracle's bytes and verifies."""
    ids, traces, cc, _ = tracegen.run_system(orc, np.random.default_rng(4), tracegen.storage_program())
    assert ids == [0, 1, 3, 4, 5, 7, 10]
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == orc.stark_prove(ids, traces, compress_challenges=cc)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg


@pytest.mark.parametrize("name", ["malloc", "storage", "poseidon_hash", "fibo_loop"])
def test_reference_prophet_programs_run_and_prove(ctx, orc, name):
    """The reference's malloc-prophet test programs (heap and write-once memory regions; `storage` = its own sstore / sload
    program, `poseidon_hash` = its poseidon-opcode program with calldata on the initial tape): GPU proof bytes = oracle's."""
    from test_oracle_stark import _reference_run

    ids, traces, cc, _ = _reference_run(orc, name)
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    assert got == orc.stark_prove(ids, traces, compress_challenges=cc)
    ok, msg = olavm_b200.verify_subsystem_proof(ids, got)
    assert ok, msg


def test_fib_loop_2p18_proof_bytes_equal_oracle(ctx, orc):
    """BASELINE configs[2] at a size the oracle finishes in about a minute: the vectorised fib-loop system
    (workload/fibloop.py; satisfying traces, all 12 tables, all 19 lookups, CPU table 2^18 rows, Program 2^19,
    Memory / RangeCheck 2^18), quotient-degree check ON: GPU proof bytes == oracle proof bytes, both verifiers accept."""
    from olavm_b200 import generation
    from workload import fibloop

    n = fibloop.bound_for_rows(18)
    ids, traces, cc, info = fibloop.fib_loop_system(n, generation.Hasher(ctx), log_n_cpu=18)
    assert info["table_log_n"][0] == 18 and ids == list(range(12))
    got = olavm_b200.prove_with_traces(ctx, ids, traces, compress_challenges=cc)
    ok, msg = olavm_b200.verify_proof(ids, got)
    assert ok, msg
    ok, msg = orc.stark_verify(ids, got)
    assert ok, msg
    ref = orc.stark_prove(ids, traces, check_degree=True, compress_challenges=cc)
    assert len(got) == len(ref) and got == ref


class _HostChallenger:
    """plonky2's Challenger (iop/challenger.rs:18-162) as the HOST application would keep it: duplex sponge over the Poseidon
    permutation, rate 8, overwrite mode, output popped from the end.  Written against the Rust, independent of the library."""

    def __init__(self, permute):
        self.permute, self.state, self.inp, self.out = permute, [0] * 12, [], []

    def _duplex(self):
        for i, x in enumerate(self.inp):
            self.state[i] = x
        self.inp = []
        self.state = [int(x) for x in self.permute(np.array(self.state, dtype=np.uint64))]
        self.out = self.state[:8]

    def observe_elements(self, xs):
        for x in xs:
            self.out = []
            self.inp.append(int(x) % P)
            if len(self.inp) == 8:
                self._duplex()

    def get_n_challenges(self, n):
        r = []
        for _ in range(n):
            if self.inp or not self.out:
                self._duplex()
            r.append(self.out.pop())
        return r

    def compact(self):
        if self.inp:
            self._duplex()
        self.out = []


def test_prove_session_with_the_hosts_own_challenger(ctx, orc):
    """ola_prove_session_*: the transcript is kept by the caller.  With plonky2's Challenger on the host side the bytes equal
    ola_prove's; every seam of prove_single_table / fri_proof shows up as a stage; a host that perturbs its transcript gets a
    different (still self-consistent) proof."""
    cmp_t, rc_t = _valid_cmp_rc(5, 6)
    single = olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t])
    log = []
    got = olavm_b200.prover.prove_with_challenger(ctx, [CMP, RC], [cmp_t, rc_t], _HostChallenger(orc.poseidon), log=log)
    assert got == single
    stages = {stage for kind, stage, table, count in log}
    assert stages >= set(range(1, 15)) - {3} or stages >= set(range(1, 15))   # every stage label occurs (3 only with compact events)
    assert any(kind == 3 for kind, *_ in log) and log[-1][0] == 4
    assert {table for _, stage, table, _ in log if stage >= 4} == {CMP, RC}
    # a different transcript (one extra absorbed element up front) still yields a proof, and a different one
    ch = _HostChallenger(orc.poseidon)
    ch.observe_elements([42])
    other = olavm_b200.prover.prove_with_challenger(ctx, [CMP, RC], [cmp_t, rc_t], ch)
    assert other != single and len(other) == len(single)
    # the context is usable afterwards
    assert olavm_b200.prove_with_traces(ctx, [CMP, RC], [cmp_t, rc_t]) == single
