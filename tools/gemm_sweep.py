"""Sweep the GEMM launch heuristics on the decoder's exact call shapes (cfg2: M = 7680 tokens, d = 768).

Each row is one GEMM call of gato_policy.py with the epilogue / operand formats it uses there; columns are the
heuristic's own choice and forced (pair, BN) variants.  CUDA-event median over `iters` launches, L2 flushed between
launches.  usage: python tools/gemm_sweep.py [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 7680
d = 768
f16, b16, f32 = torch.float16, torch.bfloat16, torch.float32
dev = "cuda"


def t(shape, dt, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dt)


flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def cases():
    x32 = t((M, d), f32)
    bias_d, bias_3d, bias_4d = t((d,), f32), t((3 * d,), f32), t((4 * d,), f32)
    a16, a16b = t((M, d), f16), t((M, d), b16)
    w_qkv, w_proj, w_fc, w_pr2 = t((d, 3 * d), f16, .02), t((d, d), f16, .02), t((d, 4 * d), f16, .02), t((4 * d, d), f16, .02)
    wb_qkv, wb_proj, wb_fc, wb_pr2 = w_qkv.to(b16), w_proj.to(b16), w_fc.to(b16), w_pr2.to(b16)
    qkv = torch.empty(M, 3 * d, device=dev, dtype=b16)
    x1 = torch.empty(M, d, device=dev, dtype=f32)
    fpre, fact, factb = (torch.empty(M, 4 * d, device=dev, dtype=dt) for dt in (b16, f16, b16))
    act4 = t((M, 4 * d), f16)
    act4b = act4.to(b16)
    dqkv = t((M, 3 * d), b16)
    out_d = torch.empty(M, d, device=dev, dtype=b16)
    dpre = torch.empty(M, 4 * d, device=dev, dtype=b16)
    g_fc, g_pr2, g_proj, g_qkv = (torch.zeros(*s, device=dev) for s in ((d, 4 * d), (4 * d, d), (d, d), (d, 3 * d)))
    E = ops
    return [
        ("fwd qkv      N2304 K768  BF16+bias", 2 * M * 3 * d * d,
         lambda: E.gemm(a16, w_qkv, b_mn=True, epilogue=E.EPI_BF16, out=qkv, bias=bias_3d)),
        ("fwd attnproj N768  K768  RESID", 2 * M * d * d,
         lambda: E.gemm(a16, w_proj, b_mn=True, epilogue=E.EPI_RESID_F32, out=x1, aux=x32, bias=bias_d)),
        ("fwd c_fc     N3072 K768  GELU x3", 2 * M * 4 * d * d,
         lambda: E.gemm(a16, w_fc, b_mn=True, epilogue=E.EPI_GELU_BF16, out=fpre, out2=fact, out3=factb, bias=bias_4d)),
        ("fwd c_proj2  N768  K3072 RESID", 2 * M * 4 * d * d,
         lambda: E.gemm(act4, w_pr2, b_mn=True, epilogue=E.EPI_RESID_F32, out=x1, aux=x32, bias=bias_d)),
        ("bwd dgelu    N3072 K768  DGELU", 2 * M * 4 * d * d,
         lambda: E.gemm(a16b, wb_pr2, epilogue=E.EPI_DGELU_BF16, out=dpre, aux=fpre)),
        ("bwd d c_fc   N768  K3072 BF16", 2 * M * 4 * d * d,
         lambda: E.gemm(act4b, wb_fc, epilogue=E.EPI_BF16, out=out_d)),
        ("bwd d proj   N768  K768  BF16", 2 * M * d * d,
         lambda: E.gemm(a16b, wb_proj, epilogue=E.EPI_BF16, out=out_d)),
        ("bwd d qkv    N768  K2304 BF16", 2 * M * 3 * d * d,
         lambda: E.gemm(dqkv, wb_qkv, epilogue=E.EPI_BF16, out=out_d)),
        ("wgrad c_proj2 3072x768  K=M", 2 * M * 4 * d * d,
         lambda: E.gemm(act4b, a16b, a_mn=True, b_mn=True, epilogue=E.EPI_F32, out=g_pr2, accumulate=True, M=4 * d, N=d, K=M)),
        ("wgrad c_fc    768x3072  K=M", 2 * M * 4 * d * d,
         lambda: E.gemm(a16b, act4b, a_mn=True, b_mn=True, epilogue=E.EPI_F32, out=g_fc, accumulate=True, M=d, N=4 * d, K=M)),
        ("wgrad proj    768x768   K=M", 2 * M * d * d,
         lambda: E.gemm(a16b, a16b, a_mn=True, b_mn=True, epilogue=E.EPI_F32, out=g_proj, accumulate=True, M=d, N=d, K=M)),
        ("wgrad qkv     768x2304  K=M", 2 * M * 3 * d * d,
         lambda: E.gemm(a16b, dqkv, a_mn=True, b_mn=True, epilogue=E.EPI_F32, out=g_qkv, accumulate=True, M=d, N=3 * d, K=M)),
    ]


VARIANTS = [("auto", {})] + [(f"p{p}bn{bn}", {"NEKO_GEMM_PAIR": str(p), "NEKO_GEMM_BN": str(bn)}) for p in (0, 1) for bn in (128, 256)]
EXTRA = [(k, dict(v.split("=") for v in k.split(","))) for k in os.environ.get("SWEEP_EXTRA", "").split(";") if k]

if __name__ == "__main__":
    if os.environ.get("SWEEP_ONE"):   # run one case a few times (ncu target): SWEEP_ONE=<row index>
        name, _fl, fn = cases()[int(os.environ["SWEEP_ONE"])]
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        print("ran", name)
        sys.exit(0)
    if os.environ.get("SWEEP_AUTO_ONLY"):
        VARIANTS = VARIANTS[:1]
    print(f"M={M}   us (TFLOP/s)      cuBLAS column: torch.matmul bf16 of the same M x N x K, plain bf16 output (no bias / GELU / residual)")
    print(f"{'call':38s}" + "".join(f"{n:>16s}" for n, _ in VARIANTS + EXTRA) + f"{'cuBLAS':>16s}")
    # the same contraction through cuBLAS (library reference, NOT on the product path): [M,K] x [K,N] bf16 -> bf16
    SHAPES = {"qkv": (M, 3 * d, d), "attnproj": (M, d, d), "c_fc ": (M, 4 * d, d), "c_proj2  N768": (M, d, 4 * d), "dgelu": (M, 4 * d, d),
              "d c_fc": (M, d, 4 * d), "d proj": (M, d, d), "d qkv": (M, d, 3 * d), "wgrad c_proj2": (4 * d, d, M), "wgrad c_fc": (d, 4 * d, M),
              "wgrad proj": (d, d, M), "wgrad qkv": (d, 3 * d, M)}

    def cublas_us(name):
        for key, (m_, n_, k_) in SHAPES.items():
            if key in name:
                a = t((m_, k_), b16)
                b = t((k_, n_), b16)
                o = torch.empty(m_, n_, device=dev, dtype=b16)
                return timeit(lambda: torch.matmul(a, b, out=o))
        return None

    for name, flops, fn in cases():
        cells = []
        for _vn, env in VARIANTS + EXTRA:
            for k in ("NEKO_GEMM_PAIR", "NEKO_GEMM_BN", "NEKO_GEMM_SPLITS", "NEKO_GEMM_NFAST", "NEKO_GEMM_DIRECT_STORE", "NEKO_GEMM_EPI16", "NEKO_GEMM_STREAMK"):
                os.environ.pop(k, None)
            os.environ.update(env)
            try:
                us = timeit(fn)
                cells.append(f"{us:7.1f} ({flops / us / 1e6:5.0f})")
            except Exception as e:  # noqa: BLE001
                cells.append("fail")
                print("   ", type(e).__name__, str(e)[:100])
        cu = cublas_us(name)
        cells.append(f"{cu:7.1f} ({flops / cu / 1e6:5.0f})" if cu else "-")
        print(f"{name:38s}" + "".join(f"{c:>16s}" for c in cells), flush=True)
