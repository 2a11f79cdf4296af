#!/usr/bin/env python3
"""Stress the thread-emulated coset-sharded prover on one GPU (race hunting): python tools/stress_sharded.py [iters]"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import olavm_b200
from workload import tracegen
from olavm_b200 import dist as odist

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = olavm_b200.Context(0)
rng = np.random.default_rng(5)
pairs = [(int(a), int(b)) for a, b in rng.integers(0, 2**32, size=(37, 2))]
cmp_t = tracegen.cmp_trace(pairs, 6)
rc_t = tracegen.rangecheck_trace([abs(a - b) for a, b in pairs])
single = olavm_b200.prove_with_traces(ctx, [3, 4], [cmp_t, rc_t])
bad = 0
for it in range(iters):
    for world in (2, 4, 8):
        try:
            proofs = odist.prove_sharded_local(0, world, [3, 4], [cmp_t, rc_t])
            if not all(p == single for p in proofs):
                bad += 1
                print("MISMATCH", it, world, [p == single for p in proofs], flush=True)
        except Exception:
            bad += 1
            print("ERROR", it, world, flush=True)
            traceback.print_exc()
print("done, failures:", bad)
