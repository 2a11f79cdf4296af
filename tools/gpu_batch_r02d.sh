#!/bin/bash
# Round-2 GPU call d: quotient kernel with pair-aligned accumulators; correctness first, then timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stark.py -m gpu -x -q -k "not fib_loop_2p18" 2>&1 | tail -3 | tee gpurun_out/r02d_pytest.txt
timeout 300 python tools/bench_prove.py --blake3 20 22 2>&1 | tee gpurun_out/r02d_prove.jsonl
