#!/bin/bash
# Round-2 GPU call 3c: generate_cpu_trace on the GPU against the oracle and the VM tables; the generation suite; timing at 2^22.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generation.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r03c_pytest.txt
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r03c_cpu_gen_timing.txt
import json, time
import numpy as np
import olavm_b200, oracle
ctx = olavm_b200.Context(0)
lib = ctx._lib
log_n = 22
n = 1 << log_n
rng = np.random.default_rng(1)
rec = rng.integers(0, 1 << 40, size=(n - 3, 66), dtype=np.uint64)
rec[:, 27] = np.left_shift(np.uint64(1), rng.integers(7, 32, size=n - 3).astype(np.uint64))
d_rec = ctx.upload(rec); d_out = ctx.alloc(94 * n)
ctx.check(lib.ola_generate_cpu_trace(ctx.handle, d_rec, n - 3, log_n, d_out, 1)); ctx.sync()
ctx.profile_begin()
ctx.check(lib.ola_generate_cpu_trace(ctx.handle, d_rec, n - 3, log_n, d_out, 1)); ctx.sync()
prof = ctx.profile_end()
t0 = time.perf_counter(); ref = oracle.generate_cpu_trace(rec, log_n); t1 = time.perf_counter()
got = ctx.download(d_out, (94, n))
print(json.dumps({"log_n": log_n, "kernel_ms": {k: round(v["ms"], 3) for k, v in prof.items()}, "bytes_moved_GB": round((66 + 94) * n * 8 / 1e9, 2),
                  "oracle_cpu_s_one_core": round(t1 - t0, 2), "equal": bool((got == ref).all())}))
PY
