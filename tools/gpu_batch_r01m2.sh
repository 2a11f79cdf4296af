mkdir -p gpurun_out
OLA_POSEIDON_UNROLLED=3 python -m pytest tests/test_gpu_parity.py -q -x -k "poseidon or hash_rows or merkle or commit" 2>&1 | tail -2
MODES="0 3" MINBS="4 5 6" ./tools/sweep_poseidon.sh 2>&1 | tee gpurun_out/sweep_poseidon2.txt | cut -c1-330
for m in 4 5 6; do echo "OLA_QUOT_MINB=$m"; OLA_QUOT_MINB=$m python tools/bench_prove.py 20 2>&1 | tail -1 | cut -c1-330; done 2>&1 | tee gpurun_out/quot_sweep2.txt
for c in 0 1; do echo "OLA_NTT_CONTIG_C4=$c"; OLA_NTT_CONTIG_C4=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline --prove-log-n 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels_ms_per_step'])"; done 2>&1 | tee gpurun_out/contig_c4.txt
