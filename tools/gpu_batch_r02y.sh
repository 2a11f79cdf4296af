#!/bin/bash
# Round-2 GPU call y (8 GPUs): the default bench line at N = 8 on the final tree.
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 \
    2>gpurun_out/r02y_bench_n8.err | tee gpurun_out/r02y_bench_n8.json | cut -c1-300
tail -3 gpurun_out/r02y_bench_n8.err
