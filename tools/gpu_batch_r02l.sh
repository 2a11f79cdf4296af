#!/bin/bash
# Round-2 GPU call l: register-round butterflies in isolation, canonical radix-2 form against the shift-twiddle form.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I olavm_b200/csrc -o /tmp/bfly16 tools/microbench/bfly16.cu 2>&1 | tail -3
timeout 300 /tmp/bfly16 2>&1 | tee gpurun_out/r02l_bfly16.txt
