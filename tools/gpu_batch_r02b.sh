#!/bin/bash
# Round-2 GPU call b: whole GPU suite (new: generation, fib-loop 2^18 byte-equality), then the bench on the valid workload.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02b_pytest.txt
timeout 900 python bench.py 2>gpurun_out/r02b_bench.err | tee gpurun_out/r02b_bench_n1.json | cut -c1-400
tail -5 gpurun_out/r02b_bench.err
