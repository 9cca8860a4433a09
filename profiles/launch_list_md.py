#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table of profiles/launches_*.md.
    python profiles/launch_list_md.py gpurun_out/launches_X.csv > profiles/launches_X.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v if unit.startswith("ms") else v * 1e3
    rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip()[:150], ms))
tot = defaultdict(lambda: [0, 0.0])
for k, ms in rows:
    tot[k][0] += 1
    tot[k][1] += ms
all_ms = sum(v[1] for v in tot.values())
ours = {k: v for k, v in tot.items() if k.startswith("kb::") or "kb::" in k.split("<")[0] or k.startswith("void kb::")}
ours_ms = sum(v[1] for v in ours.values())
setup = ("build_table", "fill_slots", "dict_init", "fill_u64", "fill_i32", "fill_memo2", "fill_f64")
jobs_ms = sum(v[1] for k, v in ours.items() if not any(s in k for s in setup))
print("| kernel | launches | total ms | share | share of ours | share of the jobs |")
print("|---|---|---|---|---|---|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    a = "%.1f%%" % (100 * ms / ours_ms) if k in ours else ""
    b = "%.1f%%" % (100 * ms / jobs_ms) if k in ours and not any(s in k for s in setup) else ""
    print("| %s | %d | %.3f | %.1f%% | %s | %s |" % (k, n, ms, 100 * ms / all_ms, a, b))
print()
print("%d launches, %.1f ms in all; this library's kernels: %.1f ms, of which inside the jobs: %.1f ms" % (len(rows), all_ms, ours_ms, jobs_ms))
