#!/bin/bash
# Round-2 GPU call 3j: the full -m gpu suite, smoke() and the default bench line (N = 1) on the final tree.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r03j_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r03j_smoke.txt
timeout 1500 python bench.py 2>gpurun_out/r03j_bench.err | tee gpurun_out/r03j_bench_n1.json | cut -c1-300
tail -3 gpurun_out/r03j_bench.err
