#!/bin/bash
# Round-2 GPU call h (8 GPUs): the full bench line at N = 8 (sharded proof over the library's own NCCL communicator,
# configs #4 and #5 legs included).
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 \
    2>gpurun_out/r02h_bench_n8.err | tee gpurun_out/r02h_bench_n8.json | cut -c1-300
tail -5 gpurun_out/r02h_bench_n8.err
