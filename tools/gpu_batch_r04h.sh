#!/bin/bash
# r04h: ola_prove_trace under the coset-sharded prover (thread-emulated ranks on one GPU)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_trace_json.py -m gpu -x -q -k "sharded" > gpurun_out/r04h_pytest.txt 2>&1
tail -12 gpurun_out/r04h_pytest.txt
