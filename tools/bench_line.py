"""stdin: bench.py output; prints the compact fields of the last JSON line (+ an optional label)."""
import json
import sys

label = " ".join(sys.argv[1:])
line = [ln for ln in sys.stdin.read().splitlines() if ln.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline", {})
print(label, d["config"]["name"], "n=%d" % d["n_gpus"], "ms/step", d["ms_per_step"], "tok/s", d["value"], "e2e", d["e2e"]["value"],
      "gemm_frac", r.get("frac"), "gemm_ms", r.get("gemm_ms_per_step"), "launches", d.get("gpu_launches"))
