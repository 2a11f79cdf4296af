#!/usr/bin/env python3
"""Prove-time benchmark of the CPU table (94 trace + 78 CTL-Z + 12 quotient columns) on a synthetic random trace
with binary filters, pipeline-parity mode (the trace does not satisfy the AIR; every kernel of the proof runs).
usage: python tools/bench_prove.py [--blake3] [log_n ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import olavm_b200
from workload import tracegen

BLAKE3 = "--blake3" in sys.argv  # C::Hasher = Blake3_256<32> (Blake3GoldilocksConfig) instead of Poseidon
logs = [int(x) for x in sys.argv[1:] if not x.startswith("--")] or [16, 18, 20]
ctx = olavm_b200.Context(0)
if BLAKE3:
    ctx.hasher = olavm_b200.BLAKE3
for lg in logs:
    rng = np.random.default_rng(lg)
    t = tracegen.cpu_random_trace(rng, lg)
    olavm_b200.prove_with_traces(ctx, [0], [t[:, : 1 << min(lg, 12)].copy()], check_quotient_degree=False)  # warm-up
    t0 = time.perf_counter()
    olavm_b200.prove_with_traces(ctx, [0], [t], check_quotient_degree=False)  # first full-size call grows the memory pool
    first = time.perf_counter() - t0
    ctx.profile_begin()
    t0 = time.perf_counter()
    proof = olavm_b200.prove_with_traces(ctx, [0], [t], check_quotient_degree=False)
    dt = time.perf_counter() - t0
    prof = ctx.profile_end()
    top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    print(json.dumps({"hasher": "blake3" if BLAKE3 else "poseidon", "log_n": lg, "prove_s": dt, "first_call_s": first, "proof_bytes": len(proof), "kernel_ms_total": sum(v["ms"] for v in prof.values()),
                      "kernels_ms": {k: round(v["ms"], 2) for k, v in top[:14]}}))
