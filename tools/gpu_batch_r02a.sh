#!/bin/bash
# First GPU call of round 2 (one `gpurun --timeout 900 -- ./tools/gpu_batch_r02a.sh`): the numbers round 1 ran out of GPU
# minutes for.  Everything lands in gpurun_out/r02a_*.
mkdir -p gpurun_out
# (1) both hashers: the whole GPU parity suite
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02a_pytest.txt
# (2) BASELINE config #4 (Merkle commit of 2^24 leaves x 64 columns), Poseidon and BLAKE3
for h in "" "--blake3"; do python tools/bench_merkle.py $h 24 64 3; done 2>&1 | tee gpurun_out/r02a_merkle_2p24x64.jsonl
# (3) CPU-table commitment and proof under both hashers, 2^20 and 2^22 rows
for h in "" "--blake3"; do python tools/bench_commit.py $h 22 94 3; python tools/bench_prove.py $h 20 22; done 2>&1 | tee gpurun_out/r02a_commit_prove.jsonl
# (4) the bench line (12-table proof under both hashers included)
python bench.py 2>gpurun_out/r02a_bench.err | tee gpurun_out/r02a_bench_n1.json | cut -c1-600
# (5) launch list of the bench command (kernel shares of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-log-n 18 > /dev/null 2>&1
tail -2 gpurun_out/r02a_launches_bench.csv
# (6) one full ncu capture of the constraint-quotient kernel (profiles/r01t_quotient_static_analysis.md, candidate 4):
#     the per-line stall page decides between the shared-memory staging and the lazy-primitive leads
ncu --set full --clock-control none --import-source on -k regex:quotient_kernel -s 2 -c 1 -o gpurun_out/r02a_quotient -f \
    python tools/bench_prove.py --blake3 20 > gpurun_out/r02a_quotient_ncu.log 2>&1
tail -2 gpurun_out/r02a_quotient_ncu.log
