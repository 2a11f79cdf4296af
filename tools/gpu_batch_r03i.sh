#!/bin/bash
# Round-2 GPU call 3i: four-lane tiles in the strided LDE pass (A/B).
mkdir -p gpurun_out
run() {
  env "$@" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or lde" 2>&1 | tail -1
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r03i_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$*', 'step_ms=%.2f e2e_ms=%.2f GB/s=%.1f frac=%.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['frac']), {a:round(b,2) for a,b in k.items()})"
}
run OLA_NTT_STRIDED_C4=0 | tee gpurun_out/r03i_ab.txt
run OLA_NTT_STRIDED_C4=1 | tee -a gpurun_out/r03i_ab.txt
