"""Which kernels of the step co-reside with the all-reduce CTAs (max-shared carveout, 148 x 128 threads)?  Each kernel is
timed alone and under a 2 ms spinner, with and without the device-wide prefer-shared cache config."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops  # noqa: E402
from neko_b200._lib import check, load  # noqa: E402

lib = load()
M, d, H, B, S = 7680, 768, 24, 32, 240
bf = torch.bfloat16
a = torch.randn(M, d, device="cuda").to(bf)
w = torch.randn(4 * d, d, device="cuda").to(bf)
out = torch.empty(M, 4 * d, device="cuda", dtype=bf)
x = torch.randn(M, d, device="cuda")
gamma, beta = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
y = torch.empty(M, d, device="cuda", dtype=torch.float16)
mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
qkv = torch.randn(B, S, 3 * d, device="cuda").to(bf)
fv = torch.zeros(B, dtype=torch.int32, device="cuda")
att = torch.empty(B, S, d, device="cuda", dtype=torch.float16)
lse = torch.empty(B, H, S, device="cuda")
big = torch.empty(M * d, device="cuda")
side = torch.cuda.Stream()

cases = {
    "gemm (pair 256x256)": lambda: ops.gemm(a, w, epilogue=ops.EPI_BF16, out=out),
    "layernorm_fwd": lambda: ops.layernorm_fwd(x, gamma, beta, 1e-5, y, mean, rstd),
    "attention_fwd": lambda: ops.attention_fwd(qkv, fv, H, S, att, lse),
    "torch fill": lambda: big.zero_(),
    "colsum": lambda: ops.colsum(out, torch.zeros(4 * d, device="cuda"), accumulate=True),
}


def timed(fn, spin):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if spin:
        with torch.cuda.stream(side):
            check(lib.neko_debug_spin(C.c_int(148), C.c_int(128), C.c_longlong(2_000_000), C.c_int(1), C.c_void_p(side.cuda_stream)), "spin")
        torch.cuda._sleep(200_000)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


for pref in (0, 1):
    check(lib.neko_prefer_shared_carveout(C.c_int(pref)), "carveout")
    print(f"device cache config: {'prefer shared' if pref else 'default'}")
    for name, fn in cases.items():
        print(f"  {name:24s} alone {timed(fn, False):8.1f} us   under the spinner {timed(fn, True):8.1f} us")
