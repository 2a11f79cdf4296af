#!/usr/bin/env python3
"""Golden vectors for the instruction ENCODING used by tests/tracegen.py's small VM.

The reference ships seven system contracts both as assembly text (assembler/test_data/asm/sys/*_asm.json, field "program")
and as the words its own assembler produced for them (assembler/test_data/bin/sys/*.json, field "bytecode").  This script
pairs every instruction line with its words -- the `main` scope first, labels resolved to word addresses the way the
assembler does: an
instruction takes two words when its last operand is an immediate / label or when it is mload / mstore, one otherwise
(core/src/program/decoder.rs:56-69) -- and writes the distinct (operands with labels replaced by their address, words)
pairs of the opcodes the VM models (at most 4 per opcode / operand shape / register choice) to
tests/golden/ola_encoding.json.

Run in the build container (needs /root/reference); the JSON is committed because the GPU box has no /root/reference.
"""
import json
import os
import re
import sys

REF = "/root/reference/assembler/test_data"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELLED = {"add", "mul", "eq", "neq", "gte", "and", "or", "xor", "not", "mov", "assert", "range", "jmp", "cjmp", "call", "ret", "end",
            "mload", "mstore", "poseidon", "tload", "tstore"}


def is_reg(a):
    return re.fullmatch(r"r\d", a) is not None


def length(op, args):
    if op in ("mload", "mstore"):
        return 2
    if not args:
        return 1
    last = args[-1]
    return 1 if (is_reg(last) or last == "psp" or last.startswith("[")) else 2


def vm_tuple(op, args):
    fix = lambda off: tuple(off) if isinstance(off, list) else off  # [offset register, factor] -> the VM's tuple form
    if op == "mstore":
        (base, off), val = args
        return ("mstore", base, fix(off), val)
    if op == "mload":
        dst, (base, off) = args
        return ("mload", dst, base, fix(off))
    return (op, *args)


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tracegen  # every pairing below is also checked against the VM's encoder, not just the sampled ones

    pairs, seen, per_contract, per_shape = [], set(), {}, {}
    for name in sorted(os.listdir(os.path.join(REF, "bin/sys"))):
        base = name[:-5]
        words = [int(w, 16) for w in json.load(open(os.path.join(REF, "bin/sys", name)))["bytecode"].split()]
        text = json.load(open(os.path.join(REF, "asm/sys", base + "_asm.json")))["program"]
        lines = [l.strip() for l in text.split("\n") if l.strip()]
        # the assembler moves the scope (function) labelled "main" to the front, the others keep their order
        # (assembler/src/relocate.rs:21-86)
        scopes, cur = [], None
        for l in lines:
            if l.endswith(":") and not l.startswith("."):
                cur = [l]
                scopes.append(cur)
            else:
                cur.append(l)
        scopes.sort(key=lambda sc: 0 if sc[0] == "main:" else 1)
        lines = [l for sc in scopes for l in sc]
        labels, pc, insts = {}, 0, []
        for l in lines:
            if l.endswith(":"):
                labels[l[:-1]] = pc
                continue
            parts = l.replace(", ", ",").split()
            op, args = parts[0], parts[1:]
            insts.append((pc, op, args))
            pc += length(op, args)
        assert pc == len(words), (base, pc, len(words))
        n = 0
        for pc, op, args in insts:
            if op not in MODELLED:
                continue
            res = []
            for a in args:
                m = re.fullmatch(r"\[(r\d)(?:,([+-]?\d+))?\]", a)
                mf = re.fullmatch(r"\[(r\d),(r\d)(?:,([+-]?\d+))?\]", a)  # [anchor, offset register(, factor = 1)]
                if mf:
                    res.append([mf.group(1), [mf.group(2), int(mf.group(3) or 1)]])
                elif m:
                    res.append([m.group(1), int(m.group(2) or 0)])
                elif is_reg(a) or a == "psp":  # psp: the prophet stack pointer special register (mov only)
                    res.append(a)
                elif re.fullmatch(r"[+-]?\d+", a):
                    res.append(int(a))
                else:
                    res.append(labels[a])
            w = words[pc:pc + length(op, args)]
            assert tracegen.ola_encode(vm_tuple(op, res)) == w, (base, pc, op, args)
            key = (op, json.dumps(res), tuple(w))
            n += 1
            if key in seen:
                continue
            seen.add(key)
            # a small fixture: at most 4 distinct examples per (opcode, operand shape, registers used)
            shape = (op, tuple(("M" if isinstance(a[1], list) else "m") if isinstance(a, list) else ("r" if isinstance(a, str) else "i") for a in res),
                     tuple(a if isinstance(a, str) else (a[0] if isinstance(a, list) else "") for a in res))
            per_shape[shape] = per_shape.get(shape, 0) + 1
            if per_shape[shape] > 4:
                continue
            pairs.append({"op": op, "args": res, "words": [hex(x) for x in w]})
        per_contract[base] = n
    out = {"source": "assembler/test_data/{asm,bin}/sys (reference's own assembler output)", "instructions_checked": per_contract,
           "note": "instructions_checked = instructions of the modelled opcodes whose words equalled tests/tracegen.py::ola_encode at "
                   "extraction time; pairs = a sample of at most 4 per opcode / operand shape / register choice",
           "pairs": pairs}
    path = os.path.join(ROOT, "tests/golden/ola_encoding.json")
    json.dump(out, open(path, "w"), indent=0)
    print(path, len(pairs), "distinct pairs from", sum(per_contract.values()), "instructions", os.path.getsize(path), "bytes")
    # the reference's prophet-free assembly test programs that only use modelled opcodes (executor/src/tests.rs runs them):
    # inputs of tests/test_oracle_stark.py::test_reference_programs_run_and_prove
    progs, prophets = {}, {}
    for name in ("bitwise", "range_check", "comparison", "fibo_recursive", "tape", "call", "memory", "mem_gep", "context_fetch"):
        d = json.load(open(os.path.join(REF, "asm", name + ".json")))
        assert not d.get("prophets")
        progs[name] = d["program"]
        tracegen.parse_ola_asm(d["program"])  # (the reference ships assembled binaries only for its system contracts, above)
    # programs whose only prophets are `cid.addr = malloc(cid.len)` / `printf(...)` (the built-ins the VM models): program + prophets
    for name in ("malloc", "mem_gep_vector", "poseidon", "poseidon_hash", "ptr_call", "storage", "storage_multi_keys", "storage_u32",
                 "fibo_loop", "printf", "global"):
        d = json.load(open(os.path.join(REF, "asm", name + ".json")))
        tracegen.parse_ola_prophets(d)
        progs[name] = d["program"]
        prophets[name] = d["prophets"]
    path = os.path.join(ROOT, "tests/golden/ola_programs.json")
    json.dump({"source": "assembler/test_data/asm/<name>.json, fields program and prophets", "programs": progs, "prophets": prophets},
              open(path, "w"), indent=0)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
