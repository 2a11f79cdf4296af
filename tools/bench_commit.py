#!/usr/bin/env python3
"""Per-kernel timing of PolynomialBatch::from_values on the GPU (CUDA events via ola_profile_*).
usage: python tools/bench_commit.py [log_n] [ncols] [reps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import olavm_b200
from olavm_b200 import PolynomialBatch

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 94
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = olavm_b200.Context(0)
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
print(json.dumps(out))
