#!/bin/bash
# Round-2 GPU call 3k: cooperative (16 lanes per node) Poseidon for the small Merkle levels: parity with the knob on, proof A/B.
mkdir -p gpurun_out
OLA_MERKLE_COOP=8192 timeout -k 5 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stark.py -m gpu -x -q -k "merkle or commit or poseidon or sharded or twelve" 2>&1 | tail -4 | tee gpurun_out/r03k_pytest.txt
rm -f gpurun_out/r03k_ab.txt
for t in 0 64 8192; do
  OLA_MERKLE_COOP=$t timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --merkle-log-l 0 --poseidon-table-log-n -1 2>gpurun_out/r03k_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['prove_all_tables']
print('OLA_MERKLE_COOP=$t', 'proof_s=%.4f'%p['seconds'], 'runs', [round(x,4) for x in p['seconds_runs']], 'merkle_level_ms', p['kernel_ms'].get('merkle_level'), 'launches', p['launches'], p['proof_sha256_16'], p['verified'])" | tee -a gpurun_out/r03k_ab.txt
done
tail -2 gpurun_out/r03k_bench.err
