#!/bin/bash
# quotient-kernel block size / barrier interval sweep on the CPU table (run on the GPU box)
lg=${1:-18}
for b in 128 256 384; do for s in 0 2 4 8 16 32; do
  echo -n "block=$b sync=$s "
  OLA_QUOT_BLOCK=$b OLA_QUOT_SYNC=$s python tools/bench_prove.py $lg | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('quotient_ms', d['kernels_ms'].get('quotient'), 'prove_s', round(d['prove_s'],4), 'first', round(d['first_call_s'],4))"
done; done
