#!/bin/bash
# Round-2 GPU call 3e: generate_prog_trace on the GPU against the oracle; the generation suite; timing at 2^23 rows.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generation.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r03e_pytest.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r03e_prog_gen_timing.txt
import json, time
import numpy as np
import olavm_b200, oracle
from olavm_b200 import generation
P = 0xFFFFFFFF00000001
ctx = olavm_b200.Context(0)
rng = np.random.default_rng(2)
m = 4000
prog_rows = np.stack([np.zeros(m, dtype=np.uint64)] * 4 + [np.arange(m, dtype=np.uint64), rng.integers(0, P, size=m, dtype=np.uint64)], axis=1)
k = (1 << 22) - 7                       # executed main lines; ~ 7.7 M fetched words as in the benchmark's run
rec = np.zeros((k, 66), dtype=np.uint64)
pcs = rng.integers(0, m - 1, size=k)
rec[:, 12] = pcs; rec[:, 25] = prog_rows[pcs, 5]; rec[:, 26] = (rng.random(k) < 0.84).astype(np.uint64); rec[:, 28] = prog_rows[pcs + 1, 5]; rec[:, 27] = np.uint64(1 << 31)
roots = np.arange(1, 9, dtype=np.uint64)
log_n = 23
generation.generate_prog_trace(ctx, rec[:1000], prog_rows, roots, 12)
ctx.profile_begin()
t0 = time.perf_counter(); got, beta = generation.generate_prog_trace(ctx, rec, prog_rows, roots, log_n); t1 = time.perf_counter()
prof = ctx.profile_end()
t2 = time.perf_counter(); ref, rbeta = oracle.generate_prog_trace(rec, prog_rows, roots, log_n); t3 = time.perf_counter()
print(json.dumps({"log_n": log_n, "fetched_words": int((got[16] == 1).sum()), "gpu_wall_s_host_buffers": round(t1 - t0, 3),
                  "kernel_ms": {k_: round(v["ms"], 2) for k_, v in prof.items()}, "oracle_cpu_s_one_core": round(t3 - t2, 2),
                  "equal": bool(beta == rbeta and (got == ref).all())}))
PY
