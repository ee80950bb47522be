"""torchrun --nproc-per-node 2 tools/dp_overlap_probe.py: does the all-reduce overlap with a chain of GEMMs?
Times (a) the GEMM chain alone, (b) the all-reduce of a 340 MB fp32 buffer alone, (c) both at once (all-reduce on a side
stream), for the p2p kernel at several CTA counts and for NCCL."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200 import dp, ops  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M, d = 7680, 768
bf = torch.bfloat16
a = torch.randn(M, d, device="cuda").to(bf)
w = torch.randn(4 * d, d, device="cuda").to(bf)
out = torch.empty(M, 4 * d, device="cuda", dtype=bf)
w2 = torch.randn(d, 4 * d, device="cuda").to(bf)
out2 = torch.empty(M, d, device="cuda", dtype=bf)
x = torch.randn(M, d, device="cuda")
gamma, beta = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
y = torch.empty(M, d, device="cuda", dtype=torch.float16)
mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
n = 85_000_000 // 64 * 64
arena = torch.randn(n, device="cuda")
side = torch.cuda.Stream()


def chain(kind):
    for _ in range(20):
        if kind == "gemm":
            ops.gemm(a, w, epilogue=ops.EPI_BF16, out=out)
            ops.gemm(out, w2, epilogue=ops.EPI_BF16, out=out2)
        else:
            ops.layernorm_fwd(x, gamma, beta, 1e-5, y, mean, rstd)


def run(label, comm, kind="gemm"):
    for _ in range(2):
        chain(kind)
        if comm:
            comm()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(True) for _ in range(4)]
    e[0].record()
    if comm:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            e[2].record()
            comm()
            e[3].record()
    chain(kind)
    e[1].record()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{label:34s} {kind} chain {e[0].elapsed_time(e[1]):7.3f} ms" + (f"   all-reduce {e[2].elapsed_time(e[3]):7.3f} ms" if comm else ""), flush=True)


for kind in ("gemm", "ln"):
    run("chain alone", None, kind)
for ctas, proto in ((148, "ce"), (148, "push"), (148, "pull"), (32, "push")):
    os.environ["NEKO_P2P_CTAS"] = str(ctas)
    os.environ["NEKO_P2P_PROTO"] = proto
    if rank == 0:
        print(f"--- protocol {proto}")
    st = dp._P2PState(arena, None)
    fn = lambda: st.all_reduce(0, n, 0.5)  # noqa: E731
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"p2p {ctas} CTAs alone: {e0.elapsed_time(e1):.3f} ms for {n * 4 / 1e6:.0f} MB", flush=True)
    for kind in ("gemm", "ln"):
        run(f"with p2p all-reduce, {ctas} CTAs", fn, kind)
fn = lambda: dist.all_reduce(arena)  # noqa: E731
for kind in ("gemm", "ln"):
    run("with NCCL all-reduce", fn, kind)
dist.destroy_process_group()
