"""cProfile of the host side of the end-to-end step (pinned batch -> forward -> backward -> loss.item()), cfg2."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neko_b200.policy import GatoPolicy  # noqa: E402
from neko_b200.tasks.synthetic import BENCH_CONFIGS, bench_batch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
c = BENCH_CONFIGS[name]


class _Tok:
    vocab_size = 50257


m = GatoPolicy(device="cuda", embed_dim=c["embed_dim"], layers=c["layers"], heads=c["heads"], dropout=0.0, resid_mid_channels=128,
               context_len=c["context_len"], text_tokenizer=_Tok())
m.transformer.drop.p = 0.0
m.use_cuda_graphs = True
m.train()
batch = [{k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in s.items()} for s in bench_batch(name)]


nxt = m.stage(batch, compute_loss=True)


def step():
    global nxt
    m.zero_grad()
    _, loss = m(nxt, compute_loss=True)
    loss.backward()
    nxt = m.stage(batch, compute_loss=True)     # batch producer: next step planned before this loss is read
    return loss.item()


for _ in range(5):
    step()
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    step()
torch.cuda.synchronize()
print(f"unprofiled: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
