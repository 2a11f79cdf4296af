#!/usr/bin/env python3
"""Extract the hot-path known-answer vectors held by the reference's own tests into tests/golden/.

Run in the build container (needs /root/reference); the JSON it writes is committed because the GPU box
has no /root/reference.  Sources:
  * plonky2/plonky2/src/hash/poseidon_goldilocks.rs:293-314  -- 4 Poseidon permutation vectors
  * core/src/util/poseidon_utils.rs:11-287                   -- inputs [0;12] and [1,0,..,0]: per-round
    states (after full rounds 1..3 of each half, state[0] after each partial round) and outputs
  * plonky2/field/src/goldilocks_field.rs:70-77, goldilocks_extensions.rs:27 -- generator constants
"""
import json
import os
import re

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0xFFFFFFFF00000001


def ints(body):
    body = re.sub(r"//[^\n]*", "", body)
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\b\d+\b", body)]


def main():
    g = open(os.path.join(REF, "plonky2/plonky2/src/hash/poseidon_goldilocks.rs")).read()
    m = re.search(r"let test_vectors12: Vec<\(\[u64; 12\], \[u64; 12\]\)> = vec!\[(.*?)\];\s*check_test_vectors", g, re.S)
    body = m.group(1).replace("neg_one", str(P - 1))
    nums = ints(body)
    assert len(nums) == 4 * 24, len(nums)
    kat = [{"input": nums[i * 24 : i * 24 + 12], "output": nums[i * 24 + 12 : i * 24 + 24]} for i in range(4)]

    u = open(os.path.join(REF, "core/src/util/poseidon_utils.rs")).read()
    rounds = {}
    for name, body in re.findall(r"pub const (POSEIDON_[A-Z0-9_]+): \[u64; \d+\] =\s*\[(.*?)\];", u, re.S):
        rounds[name] = ints(body)
    out = {
        "source": {
            "kat": "plonky2/plonky2/src/hash/poseidon_goldilocks.rs:293-314",
            "rounds": "core/src/util/poseidon_utils.rs:11-287",
        },
        "kat": kat,
        "rounds": rounds,
        "constants": {
            "order": P,
            "multiplicative_group_generator": 7,
            "power_of_two_generator": 1753635133440165772,
            "ext_power_of_two_generator": [0, 15659105665374529263],
        },
    }
    path = os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(kat), "KATs,", len(rounds), "round tables")


if __name__ == "__main__":
    main()
