#!/bin/bash
# Round-2 GPU call 3d: generate_memory_trace on the GPU against the oracle and the VM tables; the generation suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generation.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r03d_pytest.txt
