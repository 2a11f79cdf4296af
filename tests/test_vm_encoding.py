"""The small VM behind the real-run traces (tests/tracegen.py) encodes instructions the way the reference assembler does:
checked against the words the reference's own assembler produced for its seven system contracts
(tests/golden/ola_encoding.json, extracted by tools/extract_encoding_golden.py from assembler/test_data/{asm,bin}/sys)."""
import json
import os

import tracegen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ola_encoding.json")


def _to_tuple(op, args):
    fix = lambda off: tuple(off) if isinstance(off, list) else off  # [offset register, factor]: the register-scaled form
    if op == "mstore":
        (base, off), val = args
        return ("mstore", base, fix(off), val)
    if op == "mload":
        dst, (base, off) = args
        return ("mload", dst, base, fix(off))
    return (op, *args)


def test_instruction_words_match_the_reference_assembler():
    g = json.load(open(GOLDEN))
    assert sum(g["instructions_checked"].values()) > 10000  # instructions whose pairing the extractor verified
    ops = set()
    for p in g["pairs"]:
        want = [int(w, 16) for w in p["words"]]
        got = tracegen.ola_encode(_to_tuple(p["op"], p["args"]))
        assert got == want, (p, [hex(x) for x in got])
        ops.add(p["op"])
    assert ops == {"add", "mul", "eq", "neq", "gte", "and", "not", "mov", "assert", "range", "jmp", "cjmp", "call", "ret", "end",
                   "mload", "mstore", "poseidon", "tload", "tstore"}
