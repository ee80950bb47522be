"""Do small CTAs co-reside with the persistent 227 KB GEMM CTAs?  A spinning kernel (n CTAs x 128 threads, no shared memory)
is started on a side stream, then one GEMM is timed on the main stream.  If the GEMM takes its usual time while the spinner
is still resident, the two share SMs; if it takes ~the spinner's duration (or 2x its own), they do not."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops  # noqa: E402
from neko_b200._lib import check, load  # noqa: E402

lib = load()
M, d = 7680, 768
a = torch.randn(M, d, device="cuda").to(torch.bfloat16)
w = torch.randn(4 * d, d, device="cuda").to(torch.bfloat16)
out = torch.empty(M, 4 * d, device="cuda", dtype=torch.bfloat16)
g = torch.zeros(d, 4 * d, device="cuda")
act4 = torch.randn(M, 4 * d, device="cuda").to(torch.bfloat16)
side = torch.cuda.Stream()


def gemm_us(fn, spin=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if spin is not None:
        n, thr, carve = spin
        with torch.cuda.stream(side):
            check(lib.neko_debug_spin(C.c_int(n), C.c_int(thr), C.c_longlong(2_000_000), C.c_int(carve), C.c_void_p(side.cuda_stream)), "spin")
        torch.cuda._sleep(200_000)   # ~0.1 ms: let the spinner become resident first
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


cases = {"dgrad bf16 (pair 256x256)": lambda: ops.gemm(a, w, epilogue=ops.EPI_BF16, out=out),
         "wgrad f32 split-K": lambda: ops.gemm(a, act4, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=g, M=d, N=4 * d, K=M)}
for name, fn in cases.items():
    print(name)
    print(f"  alone                                   {gemm_us(fn):8.1f} us")
    for n in (16, 148):
        for thr in (128,):
            for carve in (0, 1):
                print(f"  spinner {n:3d} CTAs x {thr} thr, carveout {'max-shared' if carve else 'default   '} {gemm_us(fn, (n, thr, carve)):8.1f} us")
