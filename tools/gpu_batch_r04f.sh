#!/bin/bash
# r04f: bench.py's from_records leg at a small size (the function itself; the full bench line is the driver's)
mkdir -p gpurun_out
timeout 300 python -c "
import json, bench, olavm_b200
ctx = olavm_b200.Context(0)
ids, tr, cc, info, keep = bench.fib_workload(ctx, 16)
print(json.dumps(bench.prove_from_records(ctx, ids, tr, info)))
" > gpurun_out/r04f_from_records_small.json 2> gpurun_out/r04f.err
cat gpurun_out/r04f_from_records_small.json; tail -3 gpurun_out/r04f.err
