"""Trace JSON ingest and the `ola prove` flow: ola_trace_from_json (host code) against the records the text
was written from, and -- on the GPU -- ola_generate_traces against the oracle's generators and ola_prove_trace against the
verifiers.  The text is serde's layout of core::trace::trace::Trace (core/src/trace/trace.rs:320-342)."""
import json
import re

import numpy as np
import pytest

P = 0xFFFFFFFF00000001
_KEYS = dict(steps="step", memory="memory", rc_vals="rc_val", rc_kinds="rc_kind", bw_tags="bitwise_tag", bw_op0="bitwise_op0", bw_op1="bitwise_op1",
             bw_res="bitwise_res", cmp="cmp", psdn_inputs="poseidon_input", psdn_filters="poseidon_filter", pchunk="poseidon_chunk", storage="storage_hash",
             tape="tape", sccall="sccall", prog_rows="prog_row")


@pytest.fixture(scope="module")
def run(orc):
    from workload import trace_json as wj

    rec = wj.system_records(orc, np.random.default_rng(9))
    return rec, wj.records_to_json(rec, orc)


def test_parser_returns_the_records_the_text_was_written_from(run):
    from olavm_b200 import trace_json

    rec, text = run
    assert len(text) > 100000 and json.loads(text)["exec"][0]["register_selector"]["op0_reg_sel"]   # the nested serde shape
    t = trace_json.Trace(text)
    for key, kind in _KEYS.items():
        got = t.records(kind)
        assert got.shape == rec[key].shape and (got == rec[key]).all(), key
    assert t.records("storage_access_count") == rec["n_storage_access"]
    assert (t.records("roots").reshape(8) == rec["roots"]).all()
    # the reference's row counts: next power of two, at least 2 (2^16 RangeCheck, 3 * 2^16 -> 2^18 Bitwise, 1 for the CPU table)
    assert t.table_log_rows(0) == (len(rec["steps"]) - 1).bit_length()
    assert t.table_log_rows(2) == 18 and t.table_log_rows(4) == 16 and t.table_log_rows(9) == 1
    assert t.table_log_rows(7) == 8 and t.table_log_rows(11) == max(1, ((len(rec["prog_rows"]) + 7) // 8 - 1).bit_length())
    t.close()


def test_parser_accepts_any_key_order_whitespace_and_unknown_fields(run):
    from olavm_b200 import trace_json

    rec, text = run
    doc = json.loads(text)
    doc = {k: doc[k] for k in reversed(list(doc))}                      # builtin_program_hash before builtin_storage_hash
    doc["exec"] = [dict(reversed(list(s.items())), added_later={"x": [1, {"y": "}]"}]}) for s in doc["exec"]]
    doc["a_new_table"] = [{"v": 1.5e3, "w": None, "s": 'esc " [ { \\ '}]
    t = trace_json.Trace(json.dumps(doc, indent=1))
    for key, kind in _KEYS.items():
        assert (t.records(kind) == rec[key]).all(), key
    assert t.records("storage_access_count") == rec["n_storage_access"]


def test_parser_limits():
    from olavm_b200 import trace_json

    t = trace_json.Trace('{"exec":[],"ret":[18446744069414584320,18446744073709551615]}')
    assert t.records("step").shape == (0, 66) and t.table_log_rows(0) == 0 and t.table_log_rows(1) == 1
    t = trace_json.Trace('{"builtin_cmp":[{"op0":18446744073709551615,"op1":0,"gte":1,"abs_diff":2,"abs_diff_inv":3,"filter_looking_rc":1}]}')
    assert [int(x) for x in t.records("cmp")[0]] == [(1 << 64) - 1, 0, 1, 2, 3, 1]
    t = trace_json.Trace('{"tape":[{"is_init":true,"opcode":0,"addr":3,"value":4,"filter_looked":0},{"is_init":false,"opcode":512,"addr":3,"value":4,'
                         '"filter_looked":1}]}')
    assert t.records("tape").tolist() == [[1, 0, 3, 4, 0], [0, 512, 3, 4, 1]]
    t = trace_json.Trace('{"addr_program_hash":{"%s":[7,8,9],"%s":[]}}' % ("00" * 7 + "01" + "00" * 15 + "02" + "ff" * 8, "ab" * 32))
    assert t.records("prog_row").tolist() == [[1, 0, 2, (1 << 64) - 1, 0, 7], [1, 0, 2, (1 << 64) - 1, 1, 8], [1, 0, 2, (1 << 64) - 1, 2, 9]]


@pytest.mark.parametrize("text,why", [
    ("", "unexpected end"), ("[]", "expected '{'"), ('{"exec":[{"clk":1.5}]}', "float"), ('{"exec":[{"clk":-1}]}', "unsigned integer"),
    ('{"exec":[{"clk":18446744073709551616}]}', "64 bits"), ('{"exec":[{"regs":[1,2,3]}]}', "unexpected length"),
    ('{"exec":[{"regs":[1,2,3,4,5,6,7,8,9,10,11]}]}', "unexpected length"), ('{"exec":[{"clk":1}', "unexpected end"),
    ('{"exec":[]} x', "trailing"), ('{"addr_program_hash":{"12":[1]}}', "64 hex"), ('{"start_end_roots":[[1,2,3,4]]}', "pair"),
    ('{"tape":[{"is_init":maybe}]}', "unsigned integer or a bool"), ('{"exec":{"clk":1}}', "expected '['"),
    ('{"unknown":' + "[" * 100000 + "]" * 100000 + "}", "nesting deeper"),
])
def test_parser_rejects_what_serde_would_reject(text, why):
    from olavm_b200 import trace_json

    with pytest.raises(ValueError, match=re.escape(why)):
        trace_json.Trace(text)


@pytest.mark.gpu
def test_generate_traces_equals_the_oracle_generators_and_the_proof_verifies(ctx, orc, run):
    """The whole `ola prove` flow on the device: JSON -> records -> twelve tables generated in HBM -> proof.  Every table equals the
    oracle generator's on the same records; both verifiers accept the proof; the bytes equal ola_prove's on host copies of the
    tables (nothing in the flow depends on where the tables were generated)."""
    import olavm_b200
    from olavm_b200 import trace_json

    rec, text = run
    t = trace_json.Trace(text)
    tabs, logs, cc = trace_json.generate_traces(ctx, t)
    cols = [94, 29, 59, 6, 12, 134, 53, 48, 6, 26, 18, 40]
    try:
        host = [ctx.download(tabs[i], (cols[i], 1 << logs[i])) for i in range(12)]
    finally:
        for p in tabs:
            ctx.free(p)
    assert logs == [t.table_log_rows(i) for i in range(12)]
    ref = {0: orc.generate_cpu_trace(rec["steps"], logs[0]), 1: orc.generate_memory_trace(rec["memory"], logs[1]),
           3: orc.generate_cmp_trace(rec["cmp"]), 4: orc.generate_rc_trace(rec["rc_vals"], rec["rc_kinds"]),
           6: orc.generate_poseidon_chunk_trace(rec["pchunk"], logs[6]),
           7: orc.generate_storage_access_trace(rec["storage"][: rec["n_storage_access"]], rec["storage"][rec["n_storage_access"]:], logs[7]),
           8: orc.generate_tape_trace(rec["tape"], logs[8]), 9: orc.generate_sccall_trace(rec["sccall"], logs[9]),
           11: orc.generate_prog_chunk_trace(rec["prog_rows"], logs[11])}
    bw, beta_bw = orc.generate_bitwise_trace(rec["bw_tags"], rec["bw_op0"], rec["bw_op1"], rec["bw_res"])
    pt, beta_p = orc.generate_prog_trace(rec["steps"], rec["prog_rows"], rec["roots"], logs[10])
    ref[2], ref[10] = bw, pt
    assert cc[2] == beta_bw and cc[10] == beta_p and all(cc[i] == 0 for i in range(12) if i not in (2, 10))
    for i, r in ref.items():
        bad = [c for c in range(cols[i]) if not (host[i][c] == r[c]).all()]
        assert not bad, (i, bad)
    for i in range(12):
        assert orc.air_first_failure(i, host[i], compress_challenge=cc[i]) is None, i
    proof = trace_json.prove_trace(ctx, t)
    ok, why = olavm_b200.verify_proof(list(range(12)), proof)
    assert ok, why
    assert orc.stark_verify(list(range(12)), proof)[0]
    assert proof == olavm_b200.prove_with_traces(ctx, list(range(12)), host, compress_challenges=cc)


def _oracle_tables(orc, rec):
    """generate_traces with the oracle's generators -> ({table id: table}, {table id: compress challenge})."""
    na = rec["n_storage_access"]
    tabs = {0: orc.generate_cpu_trace(rec["steps"]), 1: orc.generate_memory_trace(rec["memory"]), 3: orc.generate_cmp_trace(rec["cmp"]),
            4: orc.generate_rc_trace(rec["rc_vals"], rec["rc_kinds"]), 6: orc.generate_poseidon_chunk_trace(rec["pchunk"]),
            7: orc.generate_storage_access_trace(rec["storage"][:na], rec["storage"][na:]), 8: orc.generate_tape_trace(rec["tape"]),
            9: orc.generate_sccall_trace(rec["sccall"]), 11: orc.generate_prog_chunk_trace(rec["prog_rows"])}
    tabs[2], beta_bw = orc.generate_bitwise_trace(rec["bw_tags"], rec["bw_op0"], rec["bw_op1"], rec["bw_res"])
    tabs[10], beta_p = orc.generate_prog_trace(rec["steps"], rec["prog_rows"], rec["roots"])
    return tabs, {2: beta_bw, 10: beta_p}


def test_fib_system_records_regenerate_its_tables(orc):
    """workload.trace_json.records_of_fib_system reads the executor records back out of the benchmark workload's tables; the
    oracle's generators rebuild from them the tables the records came from (every column that is not a free choice: the permuted
    lookup columns may fill unused table values in another order, the two beta-compressed tables use the transcript's beta), and
    what they build satisfies every AIR."""
    from workload import fibloop
    from workload import trace_json as wj

    ids, traces, cc, info = fibloop.fib_loop_system(9, orc)
    rec = wj.records_of_fib_system(traces, info)
    tabs, betas = _oracle_tables(orc, rec)
    free = {2: set(range(17, 59)), 4: {7, 8, 10, 11}, 10: {6, 7, 14, 15}, 3: set()}
    for i in (0, 1, 2, 4, 5, 7, 11, 10):
        if i == 5:
            continue
        got, want = tabs[i], traces[i]
        n = min(got.shape[1], want.shape[1])   # the workload pads some tables one power of two further than the reference
        rows = {0: info["cpu_steps"], 1: info["memory_accesses"], 2: info["bitwise_rows"]}.get(i, n)
        bad = [c for c in range(got.shape[0]) if c not in free.get(i, set()) and not (got[c, :min(rows, n)] == want[c, :min(rows, n)]).all()]
        assert not bad, (i, bad)
    for i, t in tabs.items():
        assert orc.air_first_failure(i, t, compress_challenge=betas.get(i, 0)) is None, i


def test_trace_from_records_is_the_trace_from_json(run):
    from olavm_b200 import trace_json
    from workload import trace_json as wj

    rec, text = run
    a = trace_json.Trace(text)
    b = trace_json.Trace.from_records(**{wj.REC_KIND_OF[k]: v for k, v in rec.items()})
    for kind in list(_KEYS.values()) + ["roots"]:
        assert (np.asarray(a.records(kind)) == np.asarray(b.records(kind))).all(), kind
    assert a.records("storage_access_count") == b.records("storage_access_count")
    assert [a.table_log_rows(i) for i in range(12)] == [b.table_log_rows(i) for i in range(12)]
    empty = trace_json.Trace.from_records()
    assert [empty.table_log_rows(i) for i in range(12)] == [0, 1, 18, 1, 16, 1, 1, 1, 1, 1, 1, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_prove_trace_coset_sharded_equals_the_single_gpu_proof(ctx, run, world):
    """ola_prove_trace under ola_set_comm (ranks as host threads on one GPU, olavm_b200.dist.LocalComm): every rank generates the
    twelve tables itself, runs its own Bitwise transcript thread, the proof is coset-sharded -- and every rank returns the
    single-GPU bytes."""
    import threading

    import olavm_b200
    from olavm_b200 import dist as odist
    from olavm_b200 import trace_json

    rec, text = run
    trace = trace_json.Trace(text)
    single = trace_json.prove_trace(ctx, trace)
    comm = odist.LocalComm(world)
    out, err = [None] * world, [None] * world

    def rank_main(rank):
        try:
            c = olavm_b200.Context(0)
            comm.attach(c, rank)
            out[rank] = trace_json.prove_trace(c, trace)
            c.close()
        except BaseException as e:  # noqa: B902
            err[rank] = e
            comm.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in err:
        if e is not None:
            raise e
    assert all(p == single for p in out)
    assert olavm_b200.verify_proof(list(range(12)), single)[0]


def test_parser_on_random_documents():
    """Property test (seeded): random records of every kind -> serde-shaped text with shuffled keys, random whitespace and extra
    unknown fields -> the parser returns the records."""
    import random

    from olavm_b200 import trace_json
    from workload import trace_json as wj

    rnd = random.Random(1234)
    rng = np.random.default_rng(1234)
    for trial in range(25):
        k = lambda: rnd.randrange(0, 5)
        big = lambda shape: rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)
        n_bw, n_rc, n_ps, n_st = k(), k(), k(), k()
        rec = dict(steps=big((k(), 66)), memory=big((k(), 15)), rc_vals=big(n_rc), rc_kinds=rng.integers(0, 4, size=n_rc).astype(np.uint64),
                   bw_tags=big(n_bw), bw_op0=big(n_bw), bw_op1=big(n_bw), bw_res=big(n_bw), cmp=big((k(), 6)), psdn_inputs=big((n_ps, 12)),
                   psdn_filters=rng.integers(0, 2, size=(n_ps, 4)).astype(np.uint64), pchunk=big((k(), 32)), storage=big((n_st, 38)),
                   n_storage_access=rnd.randrange(0, n_st + 1), tape=big((k(), 5)), sccall=big((k(), 24)), prog_rows=np.zeros((0, 6), dtype=np.uint64),
                   roots=big(8))
        rec["steps"][:, 11] &= np.uint64(0xFFFFFFFF)          # clk is a u32 in the Rust struct
        rec["pchunk"][:, 1] &= np.uint64(0xFFFFFFFF)
        rec["tape"][:, 0] &= np.uint64(1)                     # is_init is a bool
        progs = [(big(4), big(rnd.randrange(0, 20))) for _ in range(rnd.randrange(0, 3))]
        rows = [list(a) + [pc, w] for a, ws in progs for pc, w in enumerate(ws)]
        rec["prog_rows"] = np.array(rows, dtype=np.uint64).reshape(-1, 6)
        doc = json.loads(wj.records_to_json(rec))

        def shuffle(x):
            if isinstance(x, dict):
                items = [(key, shuffle(v)) for key, v in x.items()]
                if "addr_program_hash" not in x and not any(len(key) == 64 for key in x):   # the programs keep their file order
                    rnd.shuffle(items)
                out = dict(items)
                if rnd.random() < 0.3:
                    out["unknown_%d" % rnd.randrange(100)] = rnd.choice([None, 1.5, 'x"y \\ ]}', [1, [2, {"z": []}]], {}])
                return out
            if isinstance(x, list):
                return [shuffle(v) for v in x]
            return x

        top = shuffle({key: v for key, v in doc.items() if key != "addr_program_hash"})
        top["addr_program_hash"] = doc["addr_program_hash"]
        text = json.dumps(top, indent=rnd.choice([None, 0, 3]), separators=rnd.choice([(",", ":"), (", ", ": "), (" ,\n", " :\t")]))
        t = trace_json.Trace(text)
        for key, kind in _KEYS.items():
            got = t.records(kind)
            assert got.shape == rec[key].shape and (got == rec[key]).all(), (trial, key)
        assert t.records("storage_access_count") == rec["n_storage_access"] and (t.records("roots").reshape(8) == rec["roots"]).all()
        t.close()
