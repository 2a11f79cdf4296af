#!/bin/bash
# Round-2 GPU call 3b: generation -> prove on the device -> verify (no host round trip).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generation.py -m gpu -x -q -k "proven_where_they_lie" 2>&1 | tail -8 | tee gpurun_out/r03b_pytest.txt
