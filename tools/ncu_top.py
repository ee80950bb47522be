"""Print key metrics and the top stalled source lines of an .ncu-rep (first kernel)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
csv.field_size_limit(10**9)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "derived__lts__lts2xbar_bytes.sum.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
for d in data[:1]:
    print("kernel:", d[hdr.index("Kernel Name")][:80], "grid", d[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "")
    for k in keys:
        for i, h in enumerate(hdr):
            if h.endswith(k):
                print(f"  {k:70s} {d[i]:>16s} {units[i]}")
                break
    for i, h in enumerate(hdr):      # tensor-pipe utilisation, whatever this ncu version calls it
        if "pipe_tensor" in h and ("pct" in h or "cycles_active" in h):
            print(f"  {h:70s} {d[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[0]
e = starts[1] if len(starts) > 1 else len(rows)
h = rows[s + 1]
body = rows[s + 2:e]
isamp, isrc = h.index("# Samples"), h.index("Source")
tot = sum(int(r[isamp] or 0) for r in body) or 1
print(f"  top stalled SASS lines of {len(body)} ({tot} samples):")
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:28]:
    stalls = {hh: r[i] for i, hh in enumerate(h) if hh.startswith("stall_") and "(Not" not in hh and r[i] not in ("", "0")}
    best = sorted(stalls.items(), key=lambda kv: -float(kv[1]))[:2]
    print(f"   {100 * int(r[isamp]) / tot:5.1f}%  {r[isrc][:64]:64s} {best}")
