#!/bin/bash
# Round-2 first GPU call: counter evidence for the shipped quotient kernel + BASELINE config #4 numbers.
mkdir -p gpurun_out
# (1) full ncu capture of the shipped constraint-quotient kernel (CPU table, 2^20 rows, BLAKE3 so that the run is short)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quotient_kernel -s 2 -c 1 -o gpurun_out/r02a_quotient -f \
    python tools/bench_prove.py --blake3 20 > gpurun_out/r02a_quotient_ncu.log 2>&1
tail -2 gpurun_out/r02a_quotient_ncu.log
# (2) BASELINE config #4 (Merkle commit of 2^24 leaves x 64 columns), Poseidon and BLAKE3
for h in "" "--blake3"; do timeout 300 python tools/bench_merkle.py $h 24 64 3; done 2>&1 | tee gpurun_out/r02a_merkle_2p24x64.jsonl
# (3) CPU-table proof under both hashers, 2^20 and 2^22 rows (kernel shares)
for h in "" "--blake3"; do timeout 300 python tools/bench_prove.py $h 20 22; done 2>&1 | tee gpurun_out/r02a_prove.jsonl
