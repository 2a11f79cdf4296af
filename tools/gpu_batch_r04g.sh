#!/bin/bash
# r04g: final sanity of the shipped library on the GPU (generators, trace flow, C hosts, smoke)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trace_json.py tests/test_generation.py tests/test_c_host.py -m gpu -x -q -k "trace_json or prog_chunk or small or ola_prove_file or proof_verifies or c_host" > gpurun_out/r04g_pytest.txt 2>&1
tail -3 gpurun_out/r04g_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
