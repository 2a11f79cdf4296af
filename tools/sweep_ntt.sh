#!/bin/bash
# profiling-session helper: sweep the NTT launch knobs and print per-kernel ms of one bench step
mkdir -p gpurun_out
for T in 256 512 1024; do for G in 2 4 8; do
  OLA_NTT_THREADS=$T OLA_NTT_G=$G python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('T=$T G=$G step_ms=%.2f'%d['ms_per_step'], {a:round(b,2) for a,b in k.items()})"
done; done
