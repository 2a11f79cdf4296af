#!/bin/bash
# Round-2 GPU call 3g: ncu of the generation kernels (the HBM-bound row fill, the bitonic sort steps, the lookup walk kernels).
mkdir -p gpurun_out
cat > /tmp/gen_drive.py <<'PY'
import numpy as np
import olavm_b200
from olavm_b200 import generation
ctx = olavm_b200.Context(0)
lib = ctx._lib
rng = np.random.default_rng(1)
log_n = 22
n = 1 << log_n
rec = rng.integers(0, 1 << 40, size=(n - 3, 66), dtype=np.uint64)
rec[:, 27] = np.left_shift(np.uint64(1), rng.integers(7, 32, size=n - 3).astype(np.uint64))
d_rec = ctx.upload(rec); d_out = ctx.alloc(94 * n)
for _ in range(2):
    ctx.check(lib.ola_generate_cpu_trace(ctx.handle, d_rec, n - 3, log_n, d_out, 1)); ctx.sync()
vals = rng.integers(0, 1 << 32, size=n - 5, dtype=np.uint64); kinds = rng.integers(0, 4, size=n - 5, dtype=np.uint64)
d_v, d_k, d_rc = ctx.upload(vals), ctx.upload(kinds), ctx.alloc(12 * n)
ctx.check(lib.ola_generate_rangecheck_trace(ctx.handle, d_v, d_k, n - 5, log_n, d_rc, 1)); ctx.sync()
PY
timeout 600 ncu --set full --clock-control none -k regex:"cpu_fill_kernel|sort_tile_tail_kernel|sort_global_step_kernel|classify_kernel|match_kernel|rc_fill_kernel" -c 14 -o /tmp/r03g_gen -f env PYTHONPATH=$PWD python /tmp/gen_drive.py > gpurun_out/r03g_ncu.log 2>&1
ncu -i /tmp/r03g_gen.ncu-rep --page raw --csv > gpurun_out/r03g_gen.raw.csv 2>/dev/null
tail -2 gpurun_out/r03g_ncu.log; ls -la gpurun_out/r03g_gen.raw.csv
