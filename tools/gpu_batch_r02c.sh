#!/bin/bash
# Round-2 GPU call c: the rewritten quotient kernel (wide accumulation, chain row mapping, TMA-staged descriptors).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stark.py tests/test_gpu_blake3.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02c_pytest.txt
for h in "--blake3"; do timeout 300 python tools/bench_prove.py $h 20 22; done 2>&1 | tee gpurun_out/r02c_prove.jsonl
for mb in 3 5; do OLA_QUOT_MINB=$mb timeout 300 python tools/bench_prove.py --blake3 20 2>&1 | cut -c1-400; done | tee gpurun_out/r02c_minb.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quotient_kernel -s 2 -c 1 -o gpurun_out/r02c_quotient -f \
    python tools/bench_prove.py --blake3 20 > gpurun_out/r02c_quotient_ncu.log 2>&1
tail -2 gpurun_out/r02c_quotient_ncu.log
