"""Synthetic workload generators (test and benchmark infrastructure, not part of the product library).

`tracegen`  -- small Ola VM restating the reference executor + per-table trace generators (the spec; Python loops, slow);
`fibloop`   -- numpy-vectorised generator of the SAME tables for one loop-shaped program, fast enough for 2^22 CPU rows
               (BASELINE configs[2]); checked equal to `tracegen`'s output at small sizes (tests/test_workload.py).
"""
