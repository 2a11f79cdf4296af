#!/usr/bin/env python3
"""Per-kernel timing of PolynomialBatch::from_values on the GPU (CUDA events via ola_profile_*).
usage: python tools/bench_commit.py [--blake3] [log_n] [ncols] [reps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import olavm_b200
from olavm_b200 import PolynomialBatch

BLAKE3 = "--blake3" in sys.argv  # C::Hasher = Blake3_256<32> (Blake3GoldilocksConfig) instead of Poseidon
argv = [a for a in sys.argv if not a.startswith("--")]
log_n = int(argv[1]) if len(argv) > 1 else 20
ncols = int(argv[2]) if len(argv) > 2 else 94
reps = int(argv[3]) if len(argv) > 3 else 3
ctx = olavm_b200.Context(0)
if BLAKE3:
    ctx.hasher = olavm_b200.BLAKE3
rng = np.random.Generator(np.random.PCG64(4))
vals = rng.integers(0, 0xFFFFFFFF00000001, size=(ncols, 1 << log_n), dtype=np.uint64)
d = ctx.upload(vals)
b = PolynomialBatch.from_values(ctx, d, 3, False, 4, on_device=True, ncols=ncols, degree_log=log_n)
b.free()
ctx.profile_begin()
t0 = time.perf_counter()
for _ in range(reps):
    b = PolynomialBatch.from_values(ctx, d, 3, False, 4, on_device=True, ncols=ncols, degree_log=log_n)
    b.free()
ctx.sync()
wall = (time.perf_counter() - t0) / reps
prof = ctx.profile_end()
L = 1 << (log_n + 3)
perms = L * ((ncols + 7) // 8) + (L - 16)
out = {"log_n": log_n, "ncols": ncols, "wall_ms_per_commit": wall * 1e3,
       "kernels_ms_per_commit": {k: v["ms"] / reps for k, v in prof.items()},
       "poseidon_perms": perms}
hm = out["kernels_ms_per_commit"].get("poseidon_leaves", 0) + out["kernels_ms_per_commit"].get("merkle_level", 0)
if hm:
    out["perms_per_s"] = perms / (hm * 1e-3)
if BLAKE3:
    # algorithmic bytes: leaves 8 * ncols * L read + 32 * L written; levels 96 B per node
    k = out["kernels_ms_per_commit"]
    del out["poseidon_perms"]
    out["hasher"] = "blake3"
    out["blake3_compressions"] = L * ((ncols + 7) // 8) + (L - 16)
    if k.get("blake3_leaves"):
        out["leaf_kernel_GBps"] = (8 * ncols + 32) * L / (k["blake3_leaves"] * 1e-3) / 1e9
    if k.get("blake3_merkle_level"):
        out["level_kernels_GBps"] = 96 * (L - 16) / (k["blake3_merkle_level"] * 1e-3) / 1e9
print(json.dumps(out))
