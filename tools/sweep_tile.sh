#!/bin/bash
# profiling-session helper: A/B the tiled NTT thread mappings (OLA_NTT_VARIANT) on one bench step (per-kernel ms),
# checking each mapping against the oracle first
mkdir -p gpurun_out
run() {
  env "$@" python -m pytest tests/test_gpu_parity.py -q -x -k "tiled or lde_batch or ntt_large" 2>&1 | tail -1
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --prove-log-n ${PROVE:-0} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$*', 'step_ms=%.2f e2e_ms=%.2f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {a:round(b,2) for a,b in k.items()}, d.get('prove_all_tables',{}).get('seconds'), d.get('prove_all_tables',{}).get('kernel_ms'))"
}
for v in ${VARIANTS:-0 1 2 3}; do run OLA_NTT_VARIANT=$v; done
