#!/bin/bash
# r04e: deferred Bitwise challenge (worker thread + late table in prove_all): parity tests, then the records flow at 2^22
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trace_json.py tests/test_generation.py tests/test_c_host.py -m gpu -x -q -k "trace_json or bitwise or device or ola_prove_file or proof_verifies" > gpurun_out/r04e_pytest.txt 2>&1
tail -4 gpurun_out/r04e_pytest.txt
timeout 600 python tools/bench_prove_records.py --log-n 22 > gpurun_out/r04e_prove_records.json 2> gpurun_out/r04e_prove_records.err
tail -c 2500 gpurun_out/r04e_prove_records.json; tail -3 gpurun_out/r04e_prove_records.err
