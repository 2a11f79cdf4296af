#!/bin/bash
# Round-2 GPU call f (2 GPUs): the coset-sharded prover over the library's own NCCL communicator.
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
    2>gpurun_out/r02f_bench_n2.err | tee gpurun_out/r02f_bench_n2.json | cut -c1-300
tail -5 gpurun_out/r02f_bench_n2.err
