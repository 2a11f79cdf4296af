#!/usr/bin/env python3
"""BASELINE config #4: Merkle commit (MerkleTree::new_v2, cap height 4) of 2^log_l leaves x ncols columns resident in HBM,
column-major (the layout of a committed LDE): leaf hashing + level reduction, hashes/s and GB/s, Poseidon or BLAKE3.
usage: python tools/bench_merkle.py [--blake3] [log_l=24] [ncols=64] [reps=3]
Leaves are a device-side fill (values irrelevant to timing; parity is the tests' job): 2^24 x 64 x 8 B = 8.6 GB."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import olavm_b200
from olavm_b200 import PolynomialBatch

BLAKE3 = "--blake3" in sys.argv
argv = [a for a in sys.argv if not a.startswith("--")]
log_l = int(argv[1]) if len(argv) > 1 else 24
ncols = int(argv[2]) if len(argv) > 2 else 64
reps = int(argv[3]) if len(argv) > 3 else 3
ctx = olavm_b200.Context(0)
if BLAKE3:
    ctx.hasher = olavm_b200.BLAKE3
# a commitment of 2^(log_l - 3) rows at blowup 8 has exactly 2^log_l leaves of ncols columns: time its hashing kernels
log_n = log_l - 3
rng = np.random.Generator(np.random.PCG64(4))
vals = rng.integers(0, 0xFFFFFFFF00000001, size=(ncols, 1 << log_n), dtype=np.uint64)
d = ctx.upload(vals)
b = PolynomialBatch.from_values(ctx, d, 3, False, 4, on_device=True, ncols=ncols, degree_log=log_n)
b.free()
ctx.profile_begin()
for _ in range(reps):
    b = PolynomialBatch.from_values(ctx, d, 3, False, 4, on_device=True, ncols=ncols, degree_log=log_n)
    b.free()
ctx.sync()
prof = ctx.profile_end()
L = 1 << log_l
leaf_key, level_key = ("blake3_leaves", "blake3_merkle_level") if BLAKE3 else ("poseidon_leaves", "merkle_level")
leaf_ms = prof[leaf_key]["ms"] / reps
level_ms = prof[level_key]["ms"] / reps
calls = L * ((ncols + 7) // 8) + (L - 16)  # permutations (Poseidon) / compressions (BLAKE3, <= 128 columns)
print(json.dumps({"config": "BASELINE configs[3]: Merkle commit of 2^%d leaves x %d columns" % (log_l, ncols),
                  "hasher": "blake3" if BLAKE3 else "poseidon", "leaf_ms": leaf_ms, "levels_ms": level_ms,
                  "leaf_hashes_per_s": L / (leaf_ms * 1e-3), "hash_calls_per_s": calls / ((leaf_ms + level_ms) * 1e-3),
                  "algorithmic_GBps": ((8 * ncols + 32) * L + 96 * (L - 16)) / ((leaf_ms + level_ms) * 1e-3) / 1e9}))
