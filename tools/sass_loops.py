#!/usr/bin/env python3
"""Static SASS loop report: python tools/sass_loops.py <lib.so|cubin> <mangled-kernel-substring>
Prints, for the matching kernel, the number of SASS instructions and every loop (interval closed by a backward
branch) with its own size (instructions not inside a nested loop).  Multiply the own sizes by the trip counts known
from the source to estimate executed instructions per thread without a GPU."""
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, kernels = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in kernels.items():
        if pat not in name:
            continue
        ins = [(a, t) for a, t in ins if not t.startswith("NOP")]
        addrs = [a for a, _ in ins]
        print(f"== {name}: {len(ins)} instructions")
        loops = []
        for a, t in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addrs:
                    loops.append((tgt, a))
        loops = sorted(set(loops), key=lambda l: (l[0], -l[1]))
        def size(lo, hi):
            return sum(1 for a in addrs if lo <= a <= hi)
        for lo, hi in loops:
            inner = [(l, h) for l, h in loops if (l, h) != (lo, hi) and lo <= l and h <= hi]
            # own = total - union of maximal inner loops
            own = set(a for a in addrs if lo <= a <= hi)
            for l, h in inner:
                own -= set(a for a in addrs if l <= a <= h)
            mix = {}
            for a, t in ins:
                if a in own:
                    op = t.split()[0] if not t.startswith("@") else t.split()[1]
                    op = op.split(".")[0]
                    mix[op] = mix.get(op, 0) + 1
            top = ", ".join(f"{k}:{v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:6])
            depth = sum(1 for l, h in loops if l <= lo and hi <= h) - 1
            print(f"  {'  ' * depth}loop [{lo:#06x},{hi:#06x}] total {size(lo, hi)} own {len(own)}   {top}")


if __name__ == "__main__":
    main()
