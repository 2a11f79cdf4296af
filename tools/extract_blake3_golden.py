"""Generate tests/golden/blake3_kat.json from the OFFICIAL BLAKE3 implementation.

The reference's Blake3GoldilocksConfig (plonky2/plonky2/src/hash/blake3.rs:176,216,230) calls crate `blake3` 1.5.0
(Cargo.lock:220-221), which is not vendored in /root/reference.  The Python package `blake3` is a binding of that same
Rust crate (`blake3.__version__` recorded below), importable in the build container only -- so the vectors are
committed and this script is the record of how they were made:

    python tools/extract_blake3_golden.py

Input pattern: byte i = i % 251 (the pattern of the BLAKE3 team's own test_vectors.json), at the lengths that exercise
every structural case: empty, partial / full blocks, a full chunk, chunk + 1 byte, 2..8 chunks (parent nodes, unbalanced
trees), plus the exact leaf widths of the OlaVM tables (columns x 8 bytes) and the 96-byte challenger state.
"""
import json
import os

import blake3

LENGTHS = [0, 1, 2, 3, 7, 8, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 256, 752, 1023, 1024, 1025, 1072, 2047, 2048, 2049,
           3072, 3073, 4096, 4097, 5120, 5121, 6144, 7168, 8192, 8193, 16384, 31744]
# columns of the 12 OlaVM tables (ola_stark.rs:104-119) and the CPU table's Z / quotient / FRI leaf widths
TABLE_COLUMNS = [94, 29, 58, 6, 11, 134, 53, 26, 18, 78, 12, 32]


def main():
    vecs = []
    for n in sorted(set(LENGTHS + [8 * c for c in TABLE_COLUMNS])):
        data = bytes(i % 251 for i in range(n))
        vecs.append({"len": n, "hash": blake3.blake3(data).hexdigest()})
    out = {
        "source": "python package blake3 %s (binding of the official Rust crate); input byte i = i %% 251" % blake3.__version__,
        "known": {
            "empty": blake3.blake3(b"").hexdigest(),
            "abc": blake3.blake3(b"abc").hexdigest(),
        },
        "vectors": vecs,
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "blake3_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, len(vecs), "vectors")


if __name__ == "__main__":
    main()
