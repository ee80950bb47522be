"""torchrun --nproc-per-node N tools/dp_trace.py [cfg]: device timeline (CUDA events) of one eager data-parallel step with
the copy-engine all-reduce: when each bucket becomes ready, when its pushes land, when the reduction runs, when the broadcast
is done -- next to the start and end of backward."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from neko_b200 import dp  # noqa: E402
from neko_b200.tasks.synthetic import bench_batch  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
args = argparse.Namespace(head="rows", lean=False, no_graphs=True)
model, cfgd = bench.build_model(cfg, args, dev)
dp.broadcast_parameters(model)
no_text = not any(("text" in s and s["text"] is not None) for s in bench_batch(cfg, seed=0))
sync = dp.attach(model, bucket_bytes=int(os.environ.get("NEKO_DP_BUCKET_MB", "64")) << 20, no_text_tokens=no_text, backend="p2p")
batch = bench.to_device(bench_batch(cfg, seed=1234 + rank), dev)


def step(trace=False):
    model.zero_grad()
    _, loss = model(batch, compute_loss=True)
    if trace:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sync._p2p.trace = []
        sync._p2p._mark("backward begins", torch.cuda.current_stream())
    loss.backward()
    if trace:
        sync._p2p._mark("backward ends (comm joined)", torch.cuda.current_stream())


for _ in range(5):
    step()
torch.cuda.synchronize()
step(trace=True)
torch.cuda.synchronize()
tr = sync._p2p.trace
sync._p2p.trace = None
if rank == 0:
    t0 = tr[0][1]
    rows = sorted(((t0.elapsed_time(e), n) for n, e in tr))
    for t, n in rows:
        print(f"{t:8.3f} ms  {n}")
dist.destroy_process_group()
