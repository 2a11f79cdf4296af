#!/bin/bash
# Round-2 GPU call i (2 GPUs): correctness of the session API / sharded FRI layers on GPU 0, then the N = 2 bench with
# overlapped trace all-gathers (second NCCL communicator), with and without the side communicator.
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_stark.py tests/test_gpu_blake3.py -m gpu -x -q -k "not fib_loop_2p18" 2>&1 | tail -4 | tee gpurun_out/r02i_pytest.txt
for side in 1 0; do
OLA_NCCL_SIDE_COMM=$side timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$side bench.py --gpus 2 --steps 3 --warmup 3 \
    --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r02i_bench_n2_side$side.err | tee gpurun_out/r02i_bench_n2_side$side.json | cut -c1-200
tail -3 gpurun_out/r02i_bench_n2_side$side.err
done
