#!/bin/bash
# r04c: trace_from_records on the GPU (small), then the records flow at 2^22 rows
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trace_json.py tests/test_c_host.py -m gpu -x -q > gpurun_out/r04c_pytest.txt 2>&1
tail -4 gpurun_out/r04c_pytest.txt
timeout 900 python tools/bench_prove_records.py --log-n 22 > gpurun_out/r04c_prove_records.json 2> gpurun_out/r04c_prove_records.err
tail -c 3000 gpurun_out/r04c_prove_records.json; tail -5 gpurun_out/r04c_prove_records.err
