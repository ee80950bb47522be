"""Times neko_layernorm_bwd alone (CUDA events, L2 flushed between launches).  usage: python tools/ln_bwd_bench.py [N] [d]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 7680
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
x = torch.randn(N, d, device="cuda"); gamma = torch.ones(d, device="cuda")
dy = torch.randn(N, d, device="cuda").to(torch.bfloat16)
mean = x.mean(1); rstd = (x.var(1, unbiased=False) + 1e-5).rsqrt()
dx = torch.zeros(N, d, device="cuda"); dxb = torch.empty(N, d, device="cuda", dtype=torch.bfloat16)
dg = torch.zeros(d, device="cuda"); db = torch.zeros(d, device="cuda"); cs = torch.zeros(d, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx, dg, db, dxb, dx_colsum=cs)
for _ in range(3): run()
ts = []
for _ in range(20):
    flush.zero_()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
byt = N * d * (4 + 2 + 4 + 4 + 2)
print(f"N={N} d={d} env={ {k: v for k, v in os.environ.items() if k.startswith('NEKO_')} } median {ts[10]:.1f} us  ({byt / ts[10] / 1e3:.0f} GB/s algorithmic)")
