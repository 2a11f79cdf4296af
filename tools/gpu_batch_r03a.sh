#!/bin/bash
# Round-2 GPU call 3a: small Merkle levels with the straight-line permutation body (latency, not throughput): A/B on the proof.
mkdir -p gpurun_out
rm -f gpurun_out/r03a_ab.txt
for t in 0 4096 65536; do
  OLA_MERKLE_SMALL=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r03a_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['prove_all_tables']
print('OLA_MERKLE_SMALL=$t', 'proof_s=%.4f'%p['seconds'], 'runs', [round(x,4) for x in p['seconds_runs']], 'merkle_level_ms', p['kernel_ms'].get('merkle_level'), p['proof_sha256_16'])" | tee -a gpurun_out/r03a_ab.txt
done
tail -2 gpurun_out/r03a_bench.err
