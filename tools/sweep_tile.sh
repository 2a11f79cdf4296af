#!/bin/bash
# profiling-session helper: A/B the tiled NTT knobs on one bench step (per-kernel ms), then one ncu capture
mkdir -p gpurun_out
run() {
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --prove-log-n 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$*', 'step_ms=%.2f e2e_ms=%.2f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {a:round(b,2) for a,b in k.items()})"
}
run OLA_X=0
run OLA_NTT_COSET_MAJOR=0
run OLA_NTT_C4=1
run OLA_NTT_C4=1 OLA_NTT_COSET_MAJOR=0
