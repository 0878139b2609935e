#!/usr/bin/env python
"""Summarise ncu outputs into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/rN_launches.csv        > profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/rN_x_full.ncu-rep [regex]  > profiles/rNN_x_full.md

``launches`` reads the CSV of ``ncu --metrics gpu__time_duration.sum --clock-control none --csv`` (per-launch,
cold-cache, serialised: use the SHARE of the step, not the absolute time).  ``full`` reads an ``--set full``
report with ``ncu -i ... --page raw --csv`` and prints the metrics B200_PROFILING.md names.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"at::native::", "", name)
    return name[:110]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
    agg = collections.OrderedDict()
    tot = 0.0
    for n, v, u in rows:
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
    print(f"# ncu launch list: {path}\n\n{len(rows)} launches, {tot / 1e6:.3f} ms total (serialised, cold-cache)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |")


KEYS = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|dram__cycles_active\.avg.*|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|.*sm__pipe_tensor_cycles_active[^ ]*avg\.pct[^ ]*|"
    r".*sm__pipe_tensor_subpipe_hmma_cycles_active_realtime\.avg|sm__mem_tensor_cycles_active\.avg\.pct[^ ]*|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|"
    r"launch__block_size|launch__shared_mem_per_block_dynamic|launch__occupancy_limit_[a-z_]+|launch__waves_per_multiprocessor|"
    r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|sm__cycles_elapsed\.max|sm__cycles_active\.avg|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed\.sum|"
    r"smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active\.ratio|smsp__cycles_active\.avg)$")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines(True) if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("".join(lines))))
    hdr, units, data = rd[0], rd[1], rd[2:]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full: {path}\n")
    for row in data:
        if pattern and not re.search(pattern, row[ki]):
            continue
        print(f"## `{short(row[ki])}`  grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for h, u, v in zip(hdr, units, row):
            if KEYS.match(h) and v not in ("", "n/a"):
                if "stalled" in h:
                    try:
                        if float(v.replace(",", "")) < 0.05:
                            continue
                    except ValueError:
                        pass
                print(f"| {h} | {v} | {u} |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
