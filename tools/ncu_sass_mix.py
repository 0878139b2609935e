#!/usr/bin/env python
"""Per-opcode executed-instruction mix of one kernel from an ncu --set full --import-source report.

  python tools/ncu_sass_mix.py report.ncu-rep '<kernel-name regex>' [top_n]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', out)
for b in blocks[1:]:
    name, rest = b.split("\n", 1)
    if not re.search(pat, name):
        continue
    rd = csv.DictReader(io.StringIO(rest))
    mix = collections.Counter()
    stall = collections.Counter()
    tot = 0
    hot = []
    for r in rd:
        try:
            n = int(r["Instructions Executed"])
        except (KeyError, ValueError, TypeError):
            continue
        src = r["Source"].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2).split(".")[0] if m else src
        mix[op] += n
        tot += n
        hot.append((int(r["# Samples"] or 0), n, src))
        for k, v in r.items():
            if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
                stall[k] += int(v)
    print(f"## {name[:120]}\n\ninstructions executed (warp-level): {tot}\n")
    print("| opcode | executed | share |\n|---|---:|---:|")
    for op, n in mix.most_common(top):
        print(f"| {op} | {n} | {100 * n / tot:.1f}% |")
    s = sum(stall.values())
    print("\n| stall reason (all samples) | samples | share |\n|---|---:|---:|")
    for k, v in stall.most_common(10):
        print(f"| {k} | {v} | {100 * v / s:.1f}% |")
    print("\nhottest instructions by samples:\n")
    for smp, n, src in sorted(hot, reverse=True)[:12]:
        print(f"    {smp:6d} samples  {n:9d} exec  {src}")
    print()
    break
