"""Synthetic stand-ins for the reference's tasks (gato/tasks/*.py): same `sample_batch` output contract (list of
dicts with the keys / shapes / dtypes of control_task.py:298-324, text_task.py:47-54, caption_task.py:114-117,
vqa_task.py:92-96) without datasets or simulators, which are not available offline."""
from __future__ import annotations

from typing import List

import numpy as np
import torch


class _Base:
    kind = "control"

    def __init__(self, seed: int = 0, device=None, pin: bool = False):
        self.rs = np.random.RandomState(seed)
        self.device = device
        self.pin = pin

    def _t(self, a):
        t = torch.from_numpy(a)
        if self.device is not None:
            return t.to(self.device)
        return t.pin_memory() if self.pin else t


class SyntheticControlTask(_Base):
    """Continuous-control episodes (MuJoCo-shaped) or image control (Atari-shaped, `image_hw` set)."""
    kind = "control"

    def __init__(self, obs_dim: int = 17, act_dim: int = 6, image_hw: int = 0, n_actions: int = 4, **kw):
        super().__init__(**kw)
        self.obs_dim, self.act_dim, self.image_hw, self.n_actions = obs_dim, act_dim, image_hw, n_actions

    def tokens_per_timestep(self, patch: int = 16) -> int:
        if self.image_hw:
            return (self.image_hw // patch) ** 2 + 1 + 1
        return self.obs_dim + 1 + self.act_dim

    def sample_batch(self, n: int, max_tokens: int = 1024, **_kw) -> List[dict]:
        T = max(1, max_tokens // self.tokens_per_timestep())   # control_task.py:223
        out = []
        for _ in range(n):
            if self.image_hw:
                img = self.rs.randint(0, 256, size=(T, 3, self.image_hw, self.image_hw)).astype(np.float32)
                act = self.rs.randint(0, self.n_actions, size=(T, 1)).astype(np.int32)
                out.append({"images": self._t(img), "discrete_actions": self._t(act)})
            else:
                obs = (self.rs.standard_normal((T, self.obs_dim)) * 3).astype(np.float32)
                act = np.clip(self.rs.standard_normal((T, self.act_dim)), -1, 1).astype(np.float32)
                out.append({"continuous_obs": self._t(obs), "continuous_actions": self._t(act)})
        return out


class SyntheticTextTask(_Base):
    kind = "text"

    def __init__(self, vocab: int = 50257, **kw):
        super().__init__(**kw)
        self.vocab = vocab

    def sample_batch(self, n: int, max_tokens: int = 1024, **_kw) -> List[dict]:
        # context_len - 1 ids: the separator makes it context_len
        return [{"text": self.rs.randint(0, self.vocab, size=(max_tokens - 1,)).tolist()} for _ in range(n)]


class SyntheticCaptionTask(_Base):
    kind = "caption"

    def __init__(self, vocab: int = 50257, n_text: int = 32, hw: int = 224, **kw):
        super().__init__(**kw)
        self.vocab, self.n_text, self.hw = vocab, n_text, hw

    def sample_batch(self, n: int, **_kw) -> List[dict]:
        return [{"images": self._t(self.rs.randint(0, 256, size=(1, 3, self.hw, self.hw)).astype(np.uint8)),
                 "text": self._t(self.rs.randint(0, self.vocab, size=(self.n_text,)).astype(np.int64))} for _ in range(n)]


class SyntheticVqaTask(SyntheticCaptionTask):
    kind = "vqa"

    def __init__(self, n_text: int = 24, **kw):
        super().__init__(n_text=n_text, **kw)


def build_synthetic_tasks(name: str, seed: int = 1234, device=None, pin: bool = False):
    kw = dict(device=device, pin=pin)
    if name in ("cfg1",):
        return [SyntheticControlTask(17, 6, seed=seed, **kw)]
    if name == "cfg2":
        return [SyntheticControlTask(17, 6, seed=seed, **kw), SyntheticControlTask(11, 3, seed=seed + 1, **kw),
                SyntheticControlTask(17, 6, seed=seed + 2, **kw)]
    if name == "cfg3":
        return [SyntheticControlTask(image_hw=96, seed=seed, **kw)]
    if name == "cfg4":
        return [SyntheticTextTask(seed=seed, **kw)]
    if name == "cfg5":
        return [SyntheticTextTask(seed=seed, **kw), SyntheticCaptionTask(seed=seed + 1, **kw), SyntheticVqaTask(seed=seed + 2, **kw),
                SyntheticControlTask(17, 6, seed=seed + 3, **kw), SyntheticControlTask(image_hw=96, seed=seed + 4, **kw)]
    raise ValueError(name)


# ---------------------------------------------------------------------------------------------------------
# The five BASELINE.json configurations as (model hyper-parameters, one batch of the config's shape); bench.py's GPU arm
# draws its inputs from here (SURVEY.md section 8(d) "synthetic inputs").
# ---------------------------------------------------------------------------------------------------------
BENCH_CONFIGS = {
    "cfg1": dict(embed_dim=128, layers=3, heads=1, context_len=240, batch=4),
    "cfg2": dict(embed_dim=768, layers=6, heads=24, context_len=240, batch=32),
    "cfg3": dict(embed_dim=768, layers=6, heads=24, context_len=512, batch=32),
    "cfg4": dict(embed_dim=768, layers=6, heads=24, context_len=1024, batch=16),
    "cfg5": dict(embed_dim=768, layers=6, heads=24, context_len=1024, batch=32),
}


def bench_batch(name: str, seed: int = 1234) -> List[dict]:
    """One host-resident batch of configuration `name` (list of dicts, reference input contract)."""
    c = BENCH_CONFIGS[name]
    B, k = c["batch"], c["context_len"]
    if name == "cfg1":
        return SyntheticControlTask(17, 6, seed=seed).sample_batch(B, max_tokens=k)
    if name == "cfg2":   # halfcheetah / hopper / walker2d shapes, cycling
        tasks = [SyntheticControlTask(17, 6, seed=seed), SyntheticControlTask(11, 3, seed=seed + 1), SyntheticControlTask(17, 6, seed=seed + 2)]
        return [tasks[i % 3].sample_batch(1, max_tokens=k)[0] for i in range(B)]
    if name == "cfg3":   # Breakout-shaped frames: 84x84 gray -> 3 channels, zero-padded to 96x96 (control_task.py:380-389)
        return SyntheticControlTask(image_hw=96, seed=seed).sample_batch(B, max_tokens=k)
    if name == "cfg4":
        return SyntheticTextTask(seed=seed).sample_batch(B, max_tokens=k)
    if name == "cfg5":   # text .25 / caption .25 / vqa .25 / control .25 (half MuJoCo-shaped, half Atari-shaped)
        q = B // 4
        rest = B - 3 * q
        out = SyntheticTextTask(seed=seed).sample_batch(q, max_tokens=k)
        out += SyntheticCaptionTask(seed=seed + 1).sample_batch(q)
        out += SyntheticVqaTask(seed=seed + 2).sample_batch(q)
        out += SyntheticControlTask(17, 6, seed=seed + 3).sample_batch(rest // 2, max_tokens=k)
        out += SyntheticControlTask(image_hw=96, seed=seed + 4).sample_batch(rest - rest // 2, max_tokens=k)
        return out
    raise ValueError(name)
