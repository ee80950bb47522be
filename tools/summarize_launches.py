"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
One period (step) of the launch sequence is counted, delimited by the tokenize_embed_kernel marker."""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    rows.append((r["Kernel Name"], ns, r.get("Grid Size", "")))
# keep the launches of the last step
starts = [i for i, r in enumerate(rows) if "tokenize_embed" in r[0]]
if len(starts) >= 2:
    # one period of the launch sequence: from one step's tokenize marker up to the next one's (the image-embedding
    # and weight-cast launches that precede the marker belong to the period's tail; every step launches the same list).
    # The last full period counts; back-to-back markers (bench.py's front-end timing loop) are skipped.
    wins = [(a, b) for a, b in zip(starts[:-1], starts[1:]) if b - a > 20]
    if wins:
        rows = rows[wins[-1][0]:wins[-1][1]]
agg = OrderedDict()
for k, ns, _g in rows:
    name = k.split("(")[0]
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + ns)
total = sum(t for _, t in agg.values())
print(f"launches in one step: {len(rows)}   summed device time: {total / 1e6:.3f} ms (ncu-serialised, cold cache)")
print(f"{'kernel':60s} {'count':>6s} {'ms':>9s} {'share':>7s}")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {c:6d} {t / 1e6:9.3f} {100 * t / total:6.1f}%")

print()
print("launch sequence of the step (kernel, grid, us):")
for k, ns, g in rows:
    print(f"  {k.split('(')[0].replace('void ', '').replace('neko::', '')[:46]:46s} {g:>18s} {ns / 1e3:9.1f}")
