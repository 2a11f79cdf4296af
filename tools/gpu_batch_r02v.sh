#!/bin/bash
# Round-2 GPU call v: permuted_cols / generate_rc_trace on the GPU against the oracle's sequential walk; timing at 2^22.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generation.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r02v_pytest.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r02v_lookup_timing.txt
import json, time
import numpy as np
import olavm_b200, oracle
from olavm_b200 import generation
ctx = olavm_b200.Context(0)
for log_n in (16, 20, 22):
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    vals = rng.integers(0, 1 << 32, size=n - 5, dtype=np.uint64)
    kinds = rng.integers(0, 4, size=n - 5, dtype=np.uint64)
    generation.generate_rc_trace(ctx, vals, kinds, log_n)
    ctx.profile_begin()
    t0 = time.perf_counter(); t = generation.generate_rc_trace(ctx, vals, kinds, log_n); t1 = time.perf_counter()
    prof = ctx.profile_end()
    t2 = time.perf_counter(); ref = oracle.generate_rc_trace(vals, kinds.astype(np.uint8)); t3 = time.perf_counter()
    print(json.dumps({"log_n": log_n, "gpu_wall_s_host_buffers": round(t1 - t0, 4), "kernel_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                      "oracle_cpu_s": round(t3 - t2, 3), "equal": bool((t == ref).all())}))
PY
