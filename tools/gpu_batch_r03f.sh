#!/bin/bash
# Round-2 GPU call 3f (4 GPUs): the default bench line at N = 4 and N = 2 on the final tree (the strong-scaling curve).
mkdir -p gpurun_out
for n in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 \
      --no-cpu-baseline --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r03f_bench_n$n.err | tee gpurun_out/r03f_bench_n$n.json | cut -c1-200
  tail -2 gpurun_out/r03f_bench_n$n.err
done
