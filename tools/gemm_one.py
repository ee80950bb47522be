"""Run one GEMM shape a few times (for ncu captures): python tools/gemm_one.py M N K [a_mn b_mn]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
a_mn = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
b_mn = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
a = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
b = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out)
torch.cuda.synchronize()
