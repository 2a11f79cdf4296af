#!/bin/bash
# profiling-session helper: A/B the Poseidon body forms on a 94-column 2^20-row commitment and on the 12-table proof
mkdir -p gpurun_out
for m in ${MODES:-0 2}; do for b in ${MINBS:-5}; do
  echo "== OLA_POSEIDON_UNROLLED=$m OLA_POSEIDON_MINB=$b"
  OLA_POSEIDON_UNROLLED=$m OLA_POSEIDON_MINB=$b python tools/bench_commit.py 20 94 3 2>&1 | tail -3
done; done
