#!/bin/bash
# Round-2 GPU call x: the full -m gpu suite and the default bench line (N = 1) on the final tree, then the reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02x_pytest.txt
timeout 1500 python bench.py 2>gpurun_out/r02x_bench.err | tee gpurun_out/r02x_bench_n1.json | cut -c1-400
tail -3 gpurun_out/r02x_bench.err
