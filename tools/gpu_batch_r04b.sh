#!/bin/bash
# r04b: the five small-table generators and the Trace JSON -> prove flow on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trace_json.py -m gpu -x -q -k "trace_json or proof_verifies" > gpurun_out/r04b_pytest.txt 2>&1
tail -15 gpurun_out/r04b_pytest.txt
