#!/bin/bash
# Round-2 GPU call z: ncu --set full of the final tile kernels (raw pages exported on the box) + launch list of the bench command.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-log-n 0 --merkle-log-l 0 --poseidon-table-log-n -1"
for k in 8 9 10 11; do
  timeout 600 ncu --set full --clock-control none -k regex:tile_ --launch-skip $k -c 1 -o /tmp/r02z_tile_$k -f $CMD > gpurun_out/r02z_ncu_$k.log 2>&1
  ncu -i /tmp/r02z_tile_$k.ncu-rep --page raw --csv > gpurun_out/r02z_tile_$k.raw.csv 2>/dev/null
  tail -1 gpurun_out/r02z_ncu_$k.log | cut -c1-120
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv $CMD > gpurun_out/r02z_launches.log 2>&1
ls -la gpurun_out | head -20
