#!/bin/bash
# Round-2 GPU call w: Bitwise and Cmp table generators against the oracle; the whole generation suite again.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_generation.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02w_pytest.txt
