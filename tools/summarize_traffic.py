"""Sum dram__bytes_read/write and gpu__time_duration over the launches of an ncu --csv log -> one JSON object."""
import csv
import json
import sys

with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
per = {}
for r in csv.DictReader(lines):
    k = r["ID"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    per.setdefault(k, {})[r["Metric Name"]] = v * scale
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
ns = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
print(json.dumps({"launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": (rd + wr) / max(n, 1),
                  "gemm_time_ms_serialised": ns / 1e6,
                  "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every gemm_tcgen05_kernel launch of one step"}))
