#!/bin/bash
# Round-2 GPU call u: shorter folds of the shifts by 2^(32+r) / 2^(64+r); device-side equality of the two butterfly forms, parity, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ntt_shift.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02u_pytest.txt
run() {
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r02u_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$*', 'step_ms=%.2f e2e_ms=%.2f GB/s=%.1f frac=%.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['frac']), {a:round(b,2) for a,b in k.items()})"
}
run OLA_X=0 | tee gpurun_out/r02u_ab.txt
tail -3 gpurun_out/r02u_bench.err
