"""Data-parallel gradient equality on real GPUs (needs >= 2 devices): N ranks with different batches must end up
with the mean of the per-rank gradients (DDP semantics), and identical parameters after broadcast."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["NEKO_ROOT"])
from neko_b200 import dp
from neko_b200.policy import GatoPolicy
from oracle import gato_oracle as O
from oracle.make_golden import SMALL_CASES, small_batch
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = O.GatoConfig(**SMALL_CASES["mixed"]["cfg"])
class Tok: vocab_size = cfg.text_tokens
def make(seed):
    torch.manual_seed(seed)
    m = GatoPolicy(device=f"cuda:{rank}", embed_dim=cfg.embed_dim, layers=cfg.layers, heads=cfg.heads, dropout=0.0, resid_mid_channels=128,
                   context_len=cfg.context_len, text_tokenizer=Tok())
    m.transformer.drop.p = 0.0
    return m.eval()   # eval: deterministic patch-position bins (train mode draws them from the CPU RNG per call)
m = make(100 + rank)                 # different init per rank ...
dp.broadcast_parameters(m)           # ... made identical
ref = [torch.zeros_like(m._param_arena) for _ in range(world)]
dist.all_gather(ref, m._param_arena)
assert all(torch.equal(r, ref[0]) for r in ref)
batch = small_batch("mixed", cfg.text_tokens)
batch = batch[rank::world] if len(batch) >= world else batch   # every rank its own samples
# local gradients without synchronisation
_, loss = m(batch, compute_loss=True); loss.backward()
local = m._grad_arena.clone()
gathered = [torch.zeros_like(local) for _ in range(world)]
dist.all_gather(gathered, local)
mean = torch.stack(gathered).mean(0)
# now with the bucketed all-reduce fired from inside backward
m.zero_grad()
sync = dp.attach(m, bucket_bytes=1 << 16, backend=os.environ["NEKO_DP_BACKEND"])
assert sync.backend == os.environ["NEKO_DP_BACKEND"]
_, loss = m(batch, compute_loss=True); loss.backward()
torch.cuda.synchronize()
err = (m._grad_arena - mean).abs().max().item()
scale = mean.abs().max().item()
assert err <= 1e-5 * max(scale, 1.0) + 1e-6, (err, scale)
assert len(sync.launched) > 3 and sync.launched[0][0] == 0
# gradient accumulation: no_sync leaves local gradients, the next synced step reduces the SUM of both micro-steps
m.zero_grad()
with sync.no_sync():
    _, loss = m(batch, compute_loss=True); loss.backward()
torch.cuda.synchronize()
assert (m._grad_arena - local).abs().max().item() <= 1e-5 * max(scale, 1.0) + 1e-6
_, loss = m(batch, compute_loss=True); loss.backward()
torch.cuda.synchronize()
assert (m._grad_arena - 2 * mean).abs().max().item() <= 4e-2 * max(scale, 1.0)
# task mix without text: the text rows of embed_token are left out of the all-reduce, result unchanged
ctrl = [s for s in small_batch("mixed", cfg.text_tokens) if s.get("text") is None]
assert ctrl
m.zero_grad()
with sync.no_sync():
    _, loss = m(ctrl, compute_loss=True); loss.backward()
torch.cuda.synchronize()
local_c = m._grad_arena.clone()
sync2 = dp.attach(m, bucket_bytes=1 << 16, no_text_tokens=True, backend=os.environ["NEKO_DP_BACKEND"])
m.zero_grad()
_, loss = m(ctrl, compute_loss=True); loss.backward()
torch.cuda.synchronize()
assert (m._grad_arena - local_c).abs().max().item() <= 1e-5 * max(local_c.abs().max().item(), 1.0)   # same batch on both ranks
o = m._offs["embed_token.weight"]
assert float(m._grad_arena[o:o + cfg.text_tokens * cfg.embed_dim].abs().max()) == 0.0
assert all(hi <= o or lo >= o + cfg.text_tokens * cfg.embed_dim for lo, hi in sync2.launched)
dist.destroy_process_group()
print("dp ok", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("backend", ["p2p", "p2p-push", "p2p-pull", "nccl"])
def test_dp_gradients_match_mean_of_ranks(tmp_path, backend):
    """backend p2p: the own all-reduce over peer-mapped arenas (csrc/p2p_allreduce.cu) -- copy-engine pipeline (the default),
    or the SM-resident push / pull kernels; nccl: torch.distributed."""
    proto = {"p2p": "ce", "p2p-push": "push", "p2p-pull": "pull"}.get(backend, "ce")
    backend = backend.split("-")[0]
    script = tmp_path / "dp_worker.py"
    script.write_text(_WORKER)
    port = 29700 + (os.getpid() % 1000) + {"ce": 7, "push": 13, "pull": 19}[proto] * (backend == "p2p")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NEKO_ROOT=ROOT, NEKO_DP_BACKEND=backend, NEKO_P2P_PROTO=proto)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=150)
        assert p.returncode == 0, out.decode()[-3000:]
