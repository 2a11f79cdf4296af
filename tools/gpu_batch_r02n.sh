#!/bin/bash
# Round-2 GPU call n: ncu of the shift-form tile kernels (one bench step of config #2) + the launch list of the same command.
# (--set full reports of these kernels are large: the raw pages are exported on the box and reports over 25 MB are dropped so
# that gpurun_out/ stays under the 64 MiB that travel back.)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ntt_shift.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02n_pytest.txt
run() {
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0 2>gpurun_out/r02n_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']
print('$*', 'step_ms=%.2f e2e_ms=%.2f GB/s=%.1f frac=%.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['frac']), {a:round(b,2) for a,b in k.items()})"
}
run OLA_NTT_TILE_BRS=0 | tee gpurun_out/r02n_ab.txt
run OLA_NTT_TILE_BRS=1 | tee -a gpurun_out/r02n_ab.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0"
# tile launches per step: intt_strided, lde_strided, lde_contig (+ intt_contig when it runs as a tile kernel)
for k in 8 9 10 11; do
  timeout 600 ncu --set full --clock-control none -k regex:tile_ --launch-skip $k -c 1 -o gpurun_out/r02n_tile_$k -f $CMD > gpurun_out/r02n_ncu_$k.log 2>&1
  ncu -i gpurun_out/r02n_tile_$k.ncu-rep --page raw --csv > gpurun_out/r02n_tile_$k.raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/r02n_tile_$k.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/r02n_tile_$k.ncu-rep; fi
  tail -1 gpurun_out/r02n_ncu_$k.log | cut -c1-200
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches.csv $CMD > gpurun_out/r02n_launches.log 2>&1
ls -la gpurun_out
