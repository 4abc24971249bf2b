#!/usr/bin/env python
"""Fold an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python tools/launch_summary.py gpurun_out/launches_r01.csv > profiles/launches_r01.txt
"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
      "(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)")
print(f"{'share':>6} {'launches':>8} {'avg us':>10}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * a[1] / tot:5.1f}% {a[0]:8d} {a[1] / a[0]:10.1f}  {k[:140]}")
