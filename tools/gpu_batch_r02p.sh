#!/bin/bash
# Round-2 GPU call p: L2 bulk prefetch of the next tile in the contiguous passes (cp.async.bulk.prefetch.L2), register cap variant.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or lde or commit" 2>&1 | tail -3 | tee gpurun_out/r02p_pytest.txt
run() {
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r02p_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$*', 'step_ms=%.2f e2e_ms=%.2f GB/s=%.1f frac=%.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['frac']), {a:round(b,2) for a,b in k.items()})"
}
run OLA_NTT_PREFETCH=0 | tee gpurun_out/r02p_ab.txt
run OLA_NTT_PREFETCH=1 | tee -a gpurun_out/r02p_ab.txt
run OLA_NTT_PREFETCH=2 | tee -a gpurun_out/r02p_ab.txt
run OLA_NTT_PREFETCH=1 OLA_NTT_CONTIG_MINB=3 | tee -a gpurun_out/r02p_ab.txt
run OLA_NTT_PREFETCH=0 OLA_NTT_CONTIG_MINB=3 | tee -a gpurun_out/r02p_ab.txt
tail -3 gpurun_out/r02p_bench.err
