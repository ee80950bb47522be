"""Pins oracle/gato_oracle.py against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import gato_oracle as O
from oracle.make_golden import SCALE_MODES, SMALL_CASES, scale_batch, small_batch


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_continuous_tokenizer_known_answers():
    cfg = O.GatoConfig()
    # SURVEY.md section 8(c) anchors (probed on the reference)
    act = O.discretize(np.array([-1, -0.999999, 0, 0.9999999, 1, 5], np.float32), False, cfg)
    assert act.tolist() == [50257, 50257, 50769, 51280, 51281, 51281]
    obs = O.discretize(np.array([-300, -256, -1, -1e-9, 0, 1e-9, 1, 10, 255.9, 256, 1e6], np.float32), True, cfg)
    assert obs.tolist() == [50257, 50257, 50536, 50769, 50769, 50769, 51001, 51117, 51280, 51281, 51281]
    assert act.dtype == np.int32


def test_continuous_tokenizer_golden(golden_dir):
    g = _load(golden_dir, "tokenizer_kat.npz")
    cfg = O.GatoConfig()
    x = g["x"]
    assert np.array_equal(O.discretize(x, True, cfg), g["obs_ids"])
    assert np.array_equal(O.discretize(x, False, cfg), g["act_ids"])


def test_patch_position_bins_golden(golden_dir):
    g = _load(golden_dir, "tokenizer_kat.npz")
    for n in range(1, 41):
        assert np.array_equal(O.patch_position_indices(n, 128, False), g[f"pos_{n}"]), n
    assert O.patch_position_indices(6).tolist() == [10, 31, 52, 74, 95, 116]
    assert O.patch_position_indices(14).tolist() == [4, 13, 22, 31, 40, 49, 58, 68, 77, 86, 95, 104, 113, 122]


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_tokenize_configs_golden(golden_dir, name):
    g = _load(golden_dir, f"tok_{name}.npz")
    cfg = O.GatoConfig(context_len=O.CONFIGS[name]["context_len"])
    tb = O.tokenize(O.synth_batch(name, seed=int(g["seed"])), cfg)
    assert tb.tokens.dtype == np.int64 and tb.target_masks.dtype == np.float32
    assert np.array_equal(tb.tokens, g["tokens"])
    assert np.array_equal(tb.target_masks, g["target_masks"].astype(np.float32))
    assert np.array_equal(tb.token_masks, g["token_masks"].astype(np.float32))


def test_text_quirks():
    cfg = O.GatoConfig()
    tb = O.tokenize([{"text": [1, 2, 3]}], cfg)  # separator always appended
    assert tb.tokens.tolist() == [[1, 2, 3, 0]]
    assert tb.target_masks.tolist() == [[1, 1, 1, 0]]
    with pytest.raises(AssertionError):
        O.tokenize([{"continuous_obs": torch.zeros(3, 2), "continuous_actions": torch.zeros(4, 1)}], cfg)


def _check_against_record(g, w, out, logit_atol=2e-5, rowsum_atol=2e-3):
    assert np.array_equal(out.tokens.numpy(), g["tokens"])
    assert np.array_equal(out.target_masks.numpy(), g["target_masks"])
    assert np.array_equal(out.token_masks.numpy(), g["token_masks"])
    li, ei = g["logit_idx"], g["emb_idx"]
    emb = out.token_embeddings.detach().numpy()[ei[:, 0], ei[:, 1], ei[:, 2]]
    np.testing.assert_allclose(emb, g["emb_val"], rtol=0, atol=1e-6)
    lg = out.logits.detach().numpy()
    valid = g["token_masks"][li[:, 0], li[:, 1]] > 0   # padded rows are garbage by construction
    np.testing.assert_allclose(lg[li[:, 0], li[:, 1], li[:, 2]][valid], g["logit_val"][valid], rtol=0, atol=logit_atol)
    vrow = g["token_masks"] > 0
    np.testing.assert_allclose(lg.sum(-1)[vrow], g["logits_rowsum"][vrow], rtol=0, atol=rowsum_atol)
    assert abs(out.loss.item() - float(g["loss"])) < 1e-5
    for key in g.files:
        if key.startswith("gnone."):
            assert w[key[6:]].grad is None or float(w[key[6:]].grad.abs().max()) == 0.0
        elif key.startswith("gnorm."):
            name = key[6:]
            gr = w[name].grad.numpy().reshape(-1)
            ref_norm = float(g[key])
            assert abs(np.linalg.norm(gr.astype(np.float64)) - ref_norm) <= 1e-4 * max(ref_norm, 1e-3), name
            np.testing.assert_allclose(gr[g["gidx." + name]], g["gval." + name], rtol=1e-3, atol=1e-6, err_msg=name)


@pytest.mark.parametrize("case", list(SMALL_CASES))
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_forward_backward_golden(golden_dir, case, mode):
    g = _load(golden_dir, f"fwd_{case}_{mode}.npz")
    cfg = O.GatoConfig(**SMALL_CASES[case]["cfg"])
    w = O.make_weights(cfg, seed=3)
    for t in w.values():
        t.requires_grad_(True)
    batch = small_batch(case, cfg.text_tokens)
    torch.manual_seed(77)  # same global-RNG state as the reference run (train-mode patch bins)
    out = O.forward(w, batch, cfg, compute_loss=True, training=(mode == "train"))
    out.loss.backward()
    _check_against_record(g, w, out)


@pytest.mark.parametrize("name,mode", [(n, m) for n in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5") for m in SCALE_MODES[n]])
def test_model_scale_golden(golden_dir, name, mode):
    """The oracle at MODEL SCALE (d=768 / L=6 / H=24, V=52 305; cfg1 d=128 / L=3 / H=1) against the reference's own outputs
    on the reduced batches of oracle.make_golden.SCALE_BATCH: ids bit-exact, logits / loss / every gradient."""
    g = _load(golden_dir, f"scale_{name}_{mode}.npz")
    cfg = O.GatoConfig(**O.CONFIGS[name])
    w = O.make_weights(cfg, seed=0, perturb=False)
    for t in w.values():
        t.requires_grad_(True)
    torch.manual_seed(77)
    out = O.forward(w, scale_batch(name), cfg, compute_loss=True, training=(mode == "train"))
    out.loss.backward()
    g = {k: (g[k].astype(np.float32) if k in ("target_masks", "token_masks") else (g[k].astype(np.int64) if k == "tokens" else g[k]))
         for k in g.files}

    class _G(dict):
        files = list(g)
    _check_against_record(_G(g), w, out, logit_atol=5e-5, rowsum_atol=2e-2)
