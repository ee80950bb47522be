"""Times neko_gemm_bf16 on the decoder's GEMM shapes (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops  # noqa: E402


def bench(M, N, K, a_mn=False, b_mn=False, epi=None, iters=20):
    epi = ops.EPI_BF16 if epi is None else epi
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if epi == ops.EPI_BF16 else torch.float32)
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    for _ in range(3):
        ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=epi, out=out)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=epi, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    # cuBLAS for context (library baseline, not the product)
    a2 = a.t() if a_mn else a
    b2 = b if b_mn else b.t()
    for _ in range(3):
        torch.matmul(a2, b2)
    tc = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a2, b2)
        e1.record()
        torch.cuda.synchronize()
        tc.append(e0.elapsed_time(e1))
    tc.sort()
    fl = 2.0 * M * N * K
    return dict(M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, ms=round(t, 4), tflops=round(fl / t / 1e9, 1),
                cublas_ms=round(tc[len(tc) // 2], 4), cublas_tflops=round(fl / tc[len(tc) // 2] / 1e9, 1))


if __name__ == "__main__":
    shapes = [
        (7680, 2304, 768, False, True), (7680, 768, 768, False, True), (7680, 3072, 768, False, True), (7680, 768, 3072, False, True),
        (7680, 52352, 768, False, False), (15808, 2304, 768, False, True), (15808, 52352, 768, False, False),
        (768, 3072, 7680, True, True), (52352, 768, 7680, True, True), (7680, 768, 52352, False, True),
        (8192, 8192, 8192, False, False),
    ]
    for s in shapes:
        try:
            print(json.dumps(bench(*s)), flush=True)
        except Exception as e:  # noqa: BLE001
            print("FAIL", s, e, flush=True)
