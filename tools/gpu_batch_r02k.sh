#!/bin/bash
# Round-2 GPU call k: the full -m gpu suite on the tree after the re-entry commit (perm_values batch inversion,
# compose 4-way loads, C host asserts).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02k_pytest.txt
