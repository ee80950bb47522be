"""Host-side training logic on CPU: flag surface, *_prop batch mixing, LR schedule (reference formulas)."""
import math

import numpy as np
import pytest
import torch

from neko_b200.training.arguments import parse_args
from neko_b200.training.trainer import lr_at_step, split_batch_by_props


def test_flag_surface_and_aliases():
    a = parse_args(["--embed_dim=768", "--layers=6", "--heads=24", "-k=240", "--batch_size=32", "--text_prop", "0.25",
                    "--pad_seq", "--disable_grad_clip", "true", "--flash", "false", "-w"])
    assert (a.embed_dim, a.layers, a.heads, a.sequence_length, a.batch_size) == (768, 6, 24, 240, 32)
    assert a.pad_seq is True and a.disable_grad_clip is True and a.flash is False and a.use_wandb is True
    assert parse_args([]).dropout == 0.1 and parse_args([]).sequence_length == 1024 and parse_args([]).resid_mid_channels == 128
    assert parse_args(["--sequence_length", "512"]).sequence_length == 512
    with pytest.raises(AssertionError):
        parse_args(["--text_prop", "0.7", "--vqa_prop", "0.7"])


def test_prop_mixing_matches_reference_formula():
    # exact proportions: no remainder, no RNG
    assert split_batch_by_props(32, 0.25, 0.25, 0.25) == (8, 8, 8, 8)
    assert split_batch_by_props(512, 0.0, 0.0, 0.0) == (0, 0, 0, 512)
    # remainder goes to ONE task drawn by torch.multinomial over the residuals (trainer.py:140-154)
    torch.manual_seed(0)
    residuals = [0.3 * 10 - 3, 0.3 * 10 - 3, 0.0, 0.4 * 10 - 4]
    counts = np.zeros(4)
    for _ in range(200):
        t, c, v, k = split_batch_by_props(10, 0.33, 0.33, 0.0)
        assert t + c + v + k == 10 and v == 0
        assert (t, c, k) in ((4, 3, 3), (3, 4, 3), (3, 3, 4))
        counts += np.array([t - 3, c - 3, v, k - 3])
    assert counts[0] > 0 and counts[1] > 0 and counts[3] > 0
    # same RNG stream as the reference's single multinomial call
    torch.manual_seed(5)
    got = split_batch_by_props(7, 0.5, 0.0, 0.0)
    torch.manual_seed(5)
    idx = torch.multinomial(torch.tensor([0.5, 0.0, 0.0, 0.5]), num_samples=1).item()
    assert got == ((4, 0, 0, 3) if idx == 0 else (3, 0, 0, 4))


def test_lr_schedule_matches_reference_formula():
    kw = dict(warmup_steps=100, training_steps=1000, base_lr=1e-4, init_lr=1e-7, min_lr=1e-5)
    assert lr_at_step(0, **kw) == pytest.approx(1e-7)
    assert lr_at_step(100, **kw) == pytest.approx(1e-4)
    assert lr_at_step(50, **kw) == pytest.approx(1e-7 + (1e-4 - 1e-7) * 0.5)
    assert lr_at_step(1000, **kw) == pytest.approx(1e-5)
    mid = 1e-5 + 0.5 * (1e-4 - 1e-5) * (1 + math.cos(math.pi * 0.5))
    assert lr_at_step(550, **kw) == pytest.approx(mid)
    assert lr_at_step(700, cosine_decay=False, **kw) == pytest.approx(1e-4)


def test_synthetic_tasks_contract():
    from neko_b200.tasks import build_synthetic_tasks
    tasks = build_synthetic_tasks("cfg5", seed=3)
    kinds = {t.kind for t in tasks}
    assert kinds == {"text", "caption", "vqa", "control"}
    for t in tasks:
        b = t.sample_batch(2, max_tokens=1024)
        assert len(b) == 2 and isinstance(b[0], dict)
        if t.kind == "text":
            assert len(b[0]["text"]) == 1023
        if t.kind in ("caption", "vqa"):
            assert b[0]["images"].dtype == torch.uint8 and tuple(b[0]["images"].shape) == (1, 3, 224, 224)
