"""Is the LM-head forward GEMM (M=7680, N=52352, K=768) bound by its fp32 logits write?  Same GEMM with a 16-bit output."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops
M, N, K = 7680, 52352, 768
a = (torch.randn(M, K, device="cuda")).to(torch.float16)
b = (torch.randn(N, K, device="cuda") * 0.02).to(torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, epi, dt in (("fp32 out (the path)", ops.EPI_F32, torch.float32), ("fp16 out", ops.EPI_BF16, torch.float16)):
    out = torch.empty(M, N, device="cuda", dtype=dt)
    for _ in range(3):
        ops.gemm(a, b, epilogue=epi, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); ops.gemm(a, b, epilogue=epi, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[5]
    print(f"{name:22s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.0f} TFLOP/s  output {out.numel() * out.element_size() / 1e6:.0f} MB -> {out.numel() * out.element_size() / us / 1e3:.0f} GB/s written")
    del out
