"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, total, average, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^(void )?(hdpo::)?", "", name)
    rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (serialised, cold caches: shares matter, not absolutes)")
print(f"{'kernel':70s} {'n':>6s} {'avg us':>9s} {'total ms':>9s} {'share':>6s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:6d} {ns / c / 1e3:9.2f} {ns / 1e6:9.3f} {100 * ns / tot:5.1f}%")
