#!/usr/bin/env python3
"""Combine the cycles/iteration printed by ./pipes (on a B200) with the SASS loop bodies of the same binary:
   python tools/microbench/pipes_report.py tools/microbench/pipes gpurun_out/pipes.txt
prints, per kernel, the opcode mix of one source-level iteration (8 chains) and the resulting warp-instructions per
cycle per SM sub-partition."""
import collections
import re
import subprocess
import sys

binary, log = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
kern, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s+Function : _Z1kILi(\d+)E", line)
    if m:
        cur = int(m.group(1))
        kern[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur is not None:
        kern[cur].append((int(m.group(1), 16), m.group(2).strip()))
ITER = 2048
for line in open(log):
    m = re.match(r"KIND\s+(\d+)\s+(.*?)\s+cycles/iteration\(8 warps per SMSP\)=\s*([0-9.]+)", line)
    if not m:
        continue
    k, name, cyc = int(m.group(1)), m.group(2), float(m.group(3))
    ins = kern[k]
    loop = None
    for a, t in ins:
        mm = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if mm and int(mm.group(1), 16) < a:
            loop = (int(mm.group(1), 16), a)
    body = [t for a, t in ins if loop and loop[0] <= a <= loop[1] and not t.startswith("NOP")]
    mix = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]) for t in body)
    # unroll factor: the loop counter steps by it
    main_ops = max(mix.values()) if mix else 1
    unroll = max(1, round(main_ops / 8)) if main_ops >= 8 else 1
    per_iter = {op: v / unroll for op, v in mix.items() if v / unroll >= 0.5}
    total = sum(mix.values()) / unroll
    if cyc <= 0 or not body:
        print(f"KIND {k:2d} {name:34s} (loop optimised away by ptxas)")
        continue
    print(f"KIND {k:2d} {name:34s} cycles/iter {cyc:8.2f}  instr/iter {total:6.1f}  IPC/SMSP {8 * total / cyc:5.3f}   " +
          " ".join(f"{op}:{v:.0f}" for op, v in sorted(per_iter.items(), key=lambda kv: -kv[1])[:6]))
