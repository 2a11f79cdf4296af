#!/bin/bash
# Round-2 GPU call g: leaf-range-sharded FRI layers (thread ranks), then the full bench with the config #4 / #5 legs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stark.py tests/test_gpu_blake3.py tests/test_generation.py -m gpu -x -q -k "sharded or generate" 2>&1 | tail -4 | tee gpurun_out/r02g_pytest.txt
timeout 1200 python bench.py 2>gpurun_out/r02g_bench.err | tee gpurun_out/r02g_bench_n1.json | cut -c1-300
tail -5 gpurun_out/r02g_bench.err
