#!/bin/bash
# Round-2 GPU call q: per-instruction stall samples (source page) of the two LDE tile kernels; reports stay on the box.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0 --poseidon-table-log-n -1"
for k in 10 11; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:tile_ --launch-skip $k -c 1 -o /tmp/r02q_tile_$k -f $CMD > gpurun_out/r02q_ncu_$k.log 2>&1
  ncu -i /tmp/r02q_tile_$k.ncu-rep --page source --csv > gpurun_out/r02q_tile_$k.source.csv 2>/dev/null
  ls -la /tmp/r02q_tile_$k.ncu-rep gpurun_out/r02q_tile_$k.source.csv
done
