#!/bin/bash
# Round-2 GPU call j (8 GPUs): session / C-host tests on GPU 0, then the N = 8 bench (overlapped trace all-gathers,
# coset-sharded FRI layers, wide-accumulator compose / openings kernels).
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_stark.py tests/test_c_host.py -m gpu -x -q -k "session or c_host or sharded" 2>&1 | tail -15 | tee gpurun_out/r02j_pytest.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 \
    2>gpurun_out/r02j_bench_n8.err | tee gpurun_out/r02j_bench_n8.json | cut -c1-200
tail -3 gpurun_out/r02j_bench_n8.err
