#!/bin/bash
# r04d: the whole GPU suite at the final tree (host transcript rewritten: every proof goes through it), then the records flow at 2^22
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r04d_pytest.txt 2>&1
tail -4 gpurun_out/r04d_pytest.txt
timeout 600 python tools/bench_prove_records.py --log-n 22 > gpurun_out/r04d_prove_records.json 2> gpurun_out/r04d_prove_records.err
tail -c 2500 gpurun_out/r04d_prove_records.json; tail -3 gpurun_out/r04d_prove_records.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
