"""CPU oracle for the NEKO Gato training hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this file; the product (``neko_b200/``)
never does.  It is a functional restatement (numpy for the integer work, plain torch-CPU fp32
for the floating-point work) of what the reference computes on
``GatoPolicy.forward(inputs, compute_loss=True)``:

  reference                                   restated here
  ------------------------------------------  -----------------------------------
  gato/policy/input_tokenizers.py:5-30        ``mu_law_f32`` / ``discretize``
  gato/policy/embeddings.py:72-110            ``patch_position_indices``
  gato/policy/embeddings.py:28-61,111-131     ``image_embedding``
  gato/policy/gato_policy.py:195-432          ``tokenize`` / ``embed_and_interleave``
  gato/transformers/trajectory_gpt2.py:163-359,663-779   ``decoder``
  gato/policy/gato_policy.py:169-192          ``forward`` (LM head + masked cross entropy)

Pinning: the reference ships no golden vectors for this path, so the oracle
is pinned against the reference ITSELF: ``oracle/make_golden.py`` imports the untouched reference
in the build container (``oracle/ref_shim.py``), runs it on seeded inputs and commits the outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this file.

Weights are a plain ``dict[str, torch.Tensor]`` with the reference's ``state_dict`` key names
(SURVEY.md section 8(b)).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class GatoConfig:
    """Constructor arguments of the reference policy that shape the hot path
    (gato_policy.py:19-48)."""

    embed_dim: int = 768
    layers: int = 6
    heads: int = 24
    context_len: int = 1024
    text_tokens: int = 50257  # AutoTokenizer('gpt2').vocab_size, gato_policy.py:57-60
    continuous_tokens: int = 1024
    discrete_tokens: int = 1024
    mu: float = 100
    M: float = 256
    patch_size: int = 16
    resid_mid_channels: int = 128
    num_groups: int = 32
    position_vocab_size: int = 128
    use_pos_encoding: bool = True
    use_patch_pos_encoding: bool = True
    pad_seq: bool = False
    activation_fn: str = "gelu"  # 'geglu' adds the gate, trajectory_gpt2.py:267-276; 'gelu_new' = the tanh form of a pretrained GPT-2 config
    layer_norm_eps: float = 1e-5
    wte_rows: int = 1  # transformer.wte is a dead [1, d] table (vocab_size=1, gato_policy.py:102) unless --pretrained_lm brings GPT-2's own

    @property
    def vocab_size(self) -> int:  # gato_policy.py:63
        return self.text_tokens + self.discrete_tokens + self.continuous_tokens

    @property
    def continuous_start(self) -> int:  # gato_policy.py:66-70
        return self.text_tokens

    @property
    def discrete_start(self) -> int:
        return self.text_tokens + self.continuous_tokens


# --------------------------------------------------------------------------------------
# continuous tokenizer (input_tokenizers.py)
# --------------------------------------------------------------------------------------
def mu_law_f32(x: np.ndarray, mu: float = 100, M: float = 256) -> np.ndarray:
    """input_tokenizers.py:5-6 with every op rounded to fp32 like torch-CPU does.

    ``sign(x) * log(1 + mu*|x|) / log(1 + mu*M)``: mul, add, log, mul, true division, each an
    fp32 op on the tensor.  The logarithm is evaluated in fp64 and rounded once to fp32
    (correctly-rounded log); SURVEY.md section 7 records that this reproduces torch-CPU's bins
    on 2e7/2e7 samples, and tests/golden pins it.
    """
    x = np.asarray(x, dtype=np.float32)
    a = np.abs(x)
    t = (np.float32(mu) * a).astype(np.float32)
    t = (np.float32(1.0) + t).astype(np.float32)
    lg = np.log(t.astype(np.float64)).astype(np.float32)
    num = (np.sign(x).astype(np.float32) * lg).astype(np.float32)
    denom = np.float32(math.log(1 + mu * M))  # python double -> fp32 scalar operand
    return (num / denom).astype(np.float32)


def discretize(x: np.ndarray, use_mu_law: bool, cfg: GatoConfig) -> np.ndarray:
    """ContinuousTokenizer.encode, input_tokenizers.py:17-30.  Returns int32 token ids.

    No ``n_bins-1`` clamp exists in the reference: a clamped value of exactly 1.0 lands in bin
    ``n_bins`` = the first *discrete* token.  That is reproduced.
    """
    x = np.asarray(x, dtype=np.float32)
    if use_mu_law:
        x = mu_law_f32(x, cfg.mu, cfg.M)
    x = np.clip(x, np.float32(-1.0), np.float32(1.0)).astype(np.float32)
    x = (x + np.float32(1.0)).astype(np.float32)
    x = (x * np.float32(cfg.continuous_tokens / 2)).astype(np.float32)
    ids = np.trunc(x).astype(np.int32)  # .type(torch.int32) truncates toward zero
    return ids + np.int32(cfg.continuous_start)


# --------------------------------------------------------------------------------------
# patch position bins (embeddings.py:72-110)
# --------------------------------------------------------------------------------------
def patch_position_intervals(n: int, vocab: int = 128) -> np.ndarray:
    """embeddings.py:80-89: ``linspace(0,1,n+1)`` -> [lo,hi] pairs -> ``*vocab`` -> int32.

    torch.linspace (fp32) is the arithmetic the reference uses; it is a third-party op of the
    path, called here as-is.
    """
    ls = torch.linspace(0, 1, n + 1)
    iv = torch.stack([ls[:-1], ls[1:]]).T
    iv = (iv * vocab).to(dtype=torch.int32)
    return iv.numpy().copy()


def patch_position_indices(n: int, vocab: int = 128, training: bool = False) -> np.ndarray:
    """embeddings.py:91-100.  Eval: round-half-even of mean(lo, hi-1).  Train: one
    ``torch.randint(lo, hi)`` per row on the global CPU generator (rows first, then columns
    when the caller asks for them in that order)."""
    iv = patch_position_intervals(n, vocab)
    if training:
        return np.array(
            [int(torch.randint(low=int(lo), high=int(hi), size=())) for lo, hi in iv], dtype=np.int64
        )
    lo = iv[:, 0].astype(np.float32)
    hi = (iv[:, 1] - 1).astype(np.float32)
    mean = ((lo + hi) / np.float32(2.0)).astype(np.float32)
    return np.rint(mean).astype(np.int64)  # rint = half-to-even, like torch.round


# --------------------------------------------------------------------------------------
# tokenisation / interleave (gato_policy.py:195-432) -- integer part, numpy only
# --------------------------------------------------------------------------------------
@dataclass
class SampleTokens:
    """Per-sample result before padding (gato_policy.py:350-400)."""

    ids: np.ndarray  # int64 [T * tokens_per_timestep]
    target: np.ndarray  # float32, same length
    n_timesteps: int
    tokens_per_timestep: int
    n_obs: int  # observation tokens per timestep (position embedding is added to these)
    n_patches: int
    n_text: int
    n_cobs: int
    n_dobs: int
    n_cact: int
    n_dact: int


def _as_np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _text_matrix(text) -> np.ndarray:
    """gato_policy.py:264-273: list -> float tensor -> long; 1-D -> [1,L]; 2-D kept."""
    if isinstance(text, list):
        arr = np.asarray(text, dtype=np.float32)[None, :]  # torch.Tensor(list) is fp32
    else:
        arr = _as_np(text)
        if arr.ndim == 1:
            arr = arr[None, :]
    return arr.astype(np.int64)


def tokenize_sample(sample: dict, cfg: GatoConfig) -> SampleTokens:
    get = lambda k: sample.get(k) if sample.get(k) is not None else None  # noqa: E731
    T = None
    n_patches = n_text = n_cobs = n_dobs = n_cact = n_dact = 0
    cols_ids: List[np.ndarray] = []
    cols_tgt: List[np.ndarray] = []

    def _check_T(n):
        nonlocal T
        if T is None:
            T = n
        else:
            assert T == n, "number of timesteps must be the same for all modalities"

    text = get("text")
    text_ids = None
    if text is not None:
        text_ids = _text_matrix(text)
        T = text_ids.shape[0]  # gato_policy.py:277 assigns unconditionally (text comes first)
        n_text = text_ids.shape[1]

    img_ids = None
    if get("images") is not None or get("image_embeddings") is not None:
        if get("image_embeddings") is not None:  # takes precedence, gato_policy.py:286-287
            e = get("image_embeddings")
            n_img, n_patches = int(e.shape[0]), int(e.shape[1])
        else:
            im = get("images")
            assert im.shape[2] % cfg.patch_size == 0 and im.shape[3] % cfg.patch_size == 0, \
                "Image dimensions must be divisible by patch size"
            n_img = int(im.shape[0])
            n_patches = (int(im.shape[2]) // cfg.patch_size) * (int(im.shape[3]) // cfg.patch_size)
        _check_T(n_img)
        img_ids = np.zeros((n_img, n_patches), dtype=np.int64)

    cobs_ids = None
    if get("continuous_obs") is not None:
        cobs_ids = discretize(_as_np(get("continuous_obs")), True, cfg)
        _check_T(cobs_ids.shape[0])
        n_cobs = cobs_ids.shape[1]

    dobs_ids = None
    if get("discrete_obs") is not None:
        dobs_ids = _as_np(get("discrete_obs")) + cfg.discrete_start
        _check_T(dobs_ids.shape[0])
        n_dobs = dobs_ids.shape[1]

    cact_ids = None
    if get("continuous_actions") is not None:
        cact_ids = discretize(_as_np(get("continuous_actions")), False, cfg)
        _check_T(cact_ids.shape[0])
        n_cact = cact_ids.shape[1]

    dact_ids = None
    if get("discrete_actions") is not None:
        dact_ids = _as_np(get("discrete_actions")) + cfg.discrete_start
        _check_T(dact_ids.shape[0])
        n_dact = dact_ids.shape[1]

    assert T is not None, "sample has no modality"
    # order of gato_policy.py:355: image, text, continuous obs, discrete obs, separator, actions
    for ids, tgt in ((img_ids, 0.0), (text_ids, 1.0), (cobs_ids, 0.0), (dobs_ids, 0.0)):
        if ids is not None:
            cols_ids.append(ids.astype(np.int64))
            cols_tgt.append(np.full(ids.shape, tgt, dtype=np.float32))
    cols_ids.append(np.zeros((T, 1), dtype=np.int64))  # separator: id 0, target 0
    cols_tgt.append(np.zeros((T, 1), dtype=np.float32))
    for ids in (cact_ids, dact_ids):
        if ids is not None:
            cols_ids.append(ids.astype(np.int64))
            cols_tgt.append(np.ones(ids.shape, dtype=np.float32))
    ids = np.concatenate(cols_ids, axis=1)
    tgt = np.concatenate(cols_tgt, axis=1)
    n_obs = n_patches + n_text + n_cobs + n_dobs
    return SampleTokens(
        ids=ids.reshape(-1), target=tgt.reshape(-1), n_timesteps=T, tokens_per_timestep=ids.shape[1],
        n_obs=n_obs, n_patches=n_patches, n_text=n_text, n_cobs=n_cobs, n_dobs=n_dobs,
        n_cact=n_cact, n_dact=n_dact,
    )


@dataclass
class TokenizedBatch:
    tokens: np.ndarray  # int64 [B,S]
    target_masks: np.ndarray  # float32 [B,S]
    token_masks: np.ndarray  # float32 [B,S]
    samples: List[SampleTokens] = field(default_factory=list)

    @property
    def loss_mask(self) -> np.ndarray:
        """gato_policy.py:177-180: valid source position AND valid target at the next one."""
        return self.token_masks[:, :-1] * self.target_masks[:, 1:]


def tokenize(inputs: Sequence[dict], cfg: GatoConfig) -> TokenizedBatch:
    """Integer half of ``tokenize_input_dicts``: ids and masks, left-padded to the batch max
    (gato_policy.py:408-422), right-padded to ``context_len`` with ``pad_seq`` (:423-431)."""
    per = [tokenize_sample(s, cfg) for s in inputs]
    S = max(p.ids.shape[0] for p in per)
    width = S
    if cfg.pad_seq and cfg.context_len > S:
        width = cfg.context_len
    B = len(per)
    tokens = np.zeros((B, width), dtype=np.int64)
    target = np.zeros((B, width), dtype=np.float32)
    mask = np.zeros((B, width), dtype=np.float32)
    for b, p in enumerate(per):
        n = p.ids.shape[0]
        tokens[b, S - n:S] = p.ids
        target[b, S - n:S] = p.target
        mask[b, S - n:S] = 1.0
    return TokenizedBatch(tokens, target, mask, per)


# --------------------------------------------------------------------------------------
# floating-point half (torch CPU fp32)
# --------------------------------------------------------------------------------------
def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() / ACT2FN['gelu'] -- the erf form."""
    return F.gelu(x)


def image_embedding(images: torch.Tensor, w: Dict[str, torch.Tensor], cfg: GatoConfig,
                    row_pos: Optional[np.ndarray] = None, col_pos: Optional[np.ndarray] = None,
                    training: bool = False) -> torch.Tensor:
    """ImageEmbedding.forward (embeddings.py:28-61) + ResidualBlock_V2 (:111-131).

    images [T,3,H,W] (float 0..255 or uint8) -> [T, n_h*n_w, d].  ``row_pos``/``col_pos``
    override the position bins (train mode draws them from the CPU RNG, rows first).
    """
    p = cfg.patch_size
    T, C, H, W = images.shape
    assert H % p == 0 and W % p == 0, "Image dimensions must be divisible by patch size"
    n_h, n_w = H // p, W // p
    x = (images / 255.0 * 2) - 1
    x = x / math.sqrt(p)
    # 'b c (n_h p1) (n_w p2) -> (b n_h n_w) c p1 p2'
    x = x.reshape(T, C, n_h, p, n_w, p).permute(0, 2, 4, 1, 3, 5).reshape(T * n_h * n_w, C, p, p)
    pre = "image_embedding.patch_embedding."
    h = F.conv2d(gelu_erf(x), w[pre + "conv1.weight"], w[pre + "conv1.bias"], padding=1)
    h = F.group_norm(h, cfg.num_groups, w[pre + "gn2.weight"], w[pre + "gn2.bias"], eps=1e-5)
    h = F.conv2d(gelu_erf(h), w[pre + "conv2.weight"], w[pre + "conv2.bias"], padding=1)
    x = x + h
    x = x.reshape(T, n_h, n_w, C * p * p)
    x = F.linear(x, w["image_embedding.post_embedding_projection.weight"],
                 w["image_embedding.post_embedding_projection.bias"])
    if cfg.use_patch_pos_encoding:
        if row_pos is None:
            row_pos = patch_position_indices(n_h, cfg.position_vocab_size, training)
        if col_pos is None:
            col_pos = patch_position_indices(n_w, cfg.position_vocab_size, training)
        hp = w["image_embedding.patch_pos_encoding.height_pos_embedding.weight"][torch.as_tensor(row_pos)]
        wp = w["image_embedding.patch_pos_encoding.width_pos_embedding.weight"][torch.as_tensor(col_pos)]
        x = x + (hp[:, None, :] + wp[None, :, :])
    return x.reshape(T, n_h * n_w, -1)


def embed_and_interleave(inputs: Sequence[dict], tb: TokenizedBatch, w: Dict[str, torch.Tensor],
                         cfg: GatoConfig, training: bool = False,
                         patch_pos: Optional[List] = None) -> torch.Tensor:
    """Floating half of ``tokenize_input_dicts`` (gato_policy.py:275-431): gather rows of
    ``embed_token``, image patch embeddings, add ``pos_embed_observation[0..n_obs)`` to the
    observation block of every timestep, broadcast the separator vector, left-pad with zeros.
    ``patch_pos[b] = (row_bins, col_bins)`` pins the train-mode random bins per image sample."""
    d = cfg.embed_dim
    B, width = tb.tokens.shape
    S = max(p.ids.shape[0] for p in tb.samples)
    out = torch.zeros(B, width, d, dtype=torch.float32)
    E = w["embed_token.weight"]
    for b, (sample, st) in enumerate(zip(inputs, tb.samples)):
        T, tpt = st.n_timesteps, st.tokens_per_timestep
        ids = torch.from_numpy(st.ids.reshape(T, tpt))
        emb = E[ids]  # [T,tpt,d]; image / separator slots are overwritten below
        blocks = []
        col = 0
        if st.n_patches:
            if sample.get("image_embeddings") is not None:
                img = sample["image_embeddings"].to(torch.float32)
            else:
                rp, cp = (patch_pos[b] if patch_pos is not None and patch_pos[b] is not None else (None, None))
                img = image_embedding(sample["images"], w, cfg, rp, cp, training)
            blocks.append(img)
            col += st.n_patches
        rest = st.n_obs - st.n_patches
        if rest:
            blocks.append(emb[:, col:col + rest])
        obs = torch.cat(blocks, dim=1)
        if cfg.use_pos_encoding:
            obs = obs + w["pos_embed_observation.weight"][: st.n_obs][None]
        sep = torch.ones(T, 1, d) * w["separator_token"]
        act = emb[:, st.n_obs + 1:]
        seq = torch.cat([obs, sep, act], dim=1).reshape(T * tpt, d)
        n = T * tpt
        out[b, S - n:S] = seq
    return out


def decoder(x: torch.Tensor, token_masks: torch.Tensor, w: Dict[str, torch.Tensor], cfg: GatoConfig,
            drop: Optional[Dict] = None) -> torch.Tensor:
    """GPT2Model.forward on ``inputs_embeds``
    (trajectory_gpt2.py:663-679 mask, :322-358 block, :163-188 attention, :273-278 MLP, :779 ln_f).

    ``drop`` = None: all dropouts off.  Otherwise a dict of explicit multipliers (mask / (1 - p)) for the reference's four
    dropout sites, so a train-mode step can be replayed deterministically: ``"embd"`` [B,S,d] (self.drop, :707),
    ``("attn", i)`` [B,H,S,S] (attn_dropout on the softmax weights, :179), ``("resid_attn", i)`` [B,S,d] (resid_dropout
    after attn.c_proj, :254), ``("resid_mlp", i)`` [B,S,d] (mlp.dropout after mlp.c_proj, :278)."""
    drop = drop or {}
    B, S, d = x.shape
    if "embd" in drop:
        x = x * drop["embd"]
    H = cfg.heads
    dh = d // H
    pad_bias = ((1.0 - token_masks.to(torch.float32)) * -10000.0)[:, None, None, :]
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool))[None, None]
    neg = torch.tensor(-1e4, dtype=torch.float32)
    for i in range(cfg.layers):
        p = f"transformer.h.{i}."
        a = F.layer_norm(x, (d,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], cfg.layer_norm_eps)
        qkv = torch.addmm(w[p + "attn.c_attn.bias"], a.reshape(-1, d), w[p + "attn.c_attn.weight"]).reshape(B, S, 3 * d)
        q, k, v = qkv.split(d, dim=2)
        q = q.reshape(B, S, H, dh).permute(0, 2, 1, 3)
        k = k.reshape(B, S, H, dh).permute(0, 2, 3, 1)
        v = v.reshape(B, S, H, dh).permute(0, 2, 1, 3)
        s = torch.matmul(q, k) / (float(dh) ** 0.5)
        s = torch.where(causal, s, neg)
        s = s + pad_bias
        pr = torch.softmax(s, dim=-1)
        if ("attn", i) in drop:
            pr = pr * drop[("attn", i)]
        o = torch.matmul(pr, v).permute(0, 2, 1, 3).reshape(B, S, d)
        o = torch.addmm(w[p + "attn.c_proj.bias"], o.reshape(-1, d), w[p + "attn.c_proj.weight"]).reshape(B, S, d)
        if ("resid_attn", i) in drop:
            o = o * drop[("resid_attn", i)]
        x = o + x
        m = F.layer_norm(x, (d,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], cfg.layer_norm_eps)
        hpre = torch.addmm(w[p + "mlp.c_fc.bias"], m.reshape(-1, d), w[p + "mlp.c_fc.weight"])
        # ACT2FN[config.activation_function] (trajectory_gpt2.py:266): 'gelu' = erf form, 'gelu_new' = tanh form (pretrained GPT-2)
        hmid = F.gelu(hpre, approximate="tanh") if cfg.activation_fn == "gelu_new" else gelu_erf(hpre)
        if cfg.activation_fn == "geglu":
            hmid = hmid * F.linear(m.reshape(-1, d), w[p + "mlp.gated_layer.weight"], w[p + "mlp.gated_layer.bias"])
        m = torch.addmm(w[p + "mlp.c_proj.bias"], hmid, w[p + "mlp.c_proj.weight"]).reshape(B, S, d)
        if ("resid_mlp", i) in drop:
            m = m * drop[("resid_mlp", i)]
        x = x + m
    return F.layer_norm(x, (d,), w["transformer.ln_f.weight"], w["transformer.ln_f.bias"], cfg.layer_norm_eps)


def masked_cross_entropy(logits: torch.Tensor, tb_tokens: torch.Tensor, target_masks: torch.Tensor,
                         token_masks: torch.Tensor) -> torch.Tensor:
    """gato_policy.py:174-186: shift by one, select rows where both masks are on, mean CE."""
    V = logits.shape[-1]
    lm = (token_masks[:, :-1] * target_masks[:, 1:]).reshape(-1) > 0
    sel = logits[:, :-1, :].reshape(-1, V)[lm]
    tgt = tb_tokens[:, 1:].reshape(-1)[lm]
    return F.cross_entropy(sel, tgt)


@dataclass
class OracleOutput:
    token_embeddings: torch.Tensor
    tokens: torch.Tensor
    target_masks: torch.Tensor
    token_masks: torch.Tensor
    hidden: torch.Tensor
    logits: torch.Tensor
    loss: Optional[torch.Tensor]


def forward(w: Dict[str, torch.Tensor], inputs: Sequence[dict], cfg: GatoConfig, compute_loss: bool = True,
            training: bool = False, patch_pos: Optional[List] = None, drop: Optional[Dict] = None) -> OracleOutput:
    """GatoPolicy.forward(inputs, compute_loss) (gato_policy.py:156-192); dropout off unless explicit multipliers are
    given in ``drop`` (see ``decoder``)."""
    tb = tokenize(inputs, cfg)
    emb = embed_and_interleave(inputs, tb, w, cfg, training, patch_pos)
    tokens = torch.from_numpy(tb.tokens)
    tmask = torch.from_numpy(tb.target_masks)
    mask = torch.from_numpy(tb.token_masks)
    hid = decoder(emb, mask, w, cfg, drop)
    logits = F.linear(hid, w["predict_token.weight"])
    loss = masked_cross_entropy(logits, tokens, tmask, mask) if compute_loss else None
    return OracleOutput(emb, tokens, tmask, mask, hid, logits, loss)


# --------------------------------------------------------------------------------------
# inference loops (greedy): gato_policy.py:444-478 predict_text, :557-616 predict_control
# --------------------------------------------------------------------------------------
def generate(w: Dict[str, torch.Tensor], emb: torch.Tensor, mask: torch.Tensor, cfg: GatoConfig, n_tokens: int, lo: int, hi: int):
    """Greedy continuation on embeddings (gato_policy.py:452-476 / :586-605): last-position logits restricted to
    ids [lo, hi] -> argmax -> embed_token row appended -> context trimmed to context_len.  Returns
    (list of restricted logit rows, list of picked absolute ids)."""
    rows, picked = [], []
    for _ in range(n_tokens):
        hid = decoder(emb, mask, w, cfg)
        row = F.linear(hid[0, -1], w["predict_token.weight"])[lo:hi + 1]
        rows.append(row)
        tok = int(torch.argmax(row)) + lo
        picked.append(tok)
        emb = torch.cat([emb, w["embed_token.weight"][tok].reshape(1, 1, -1)], dim=1)[:, -cfg.context_len:, :]
        mask = torch.cat([mask, torch.ones(mask.shape[0], 1)], dim=1)[:, -cfg.context_len:]
    return rows, picked


def predict_text(w: Dict[str, torch.Tensor], batch_dict: dict, cfg: GatoConfig, max_length: int = 20):
    tb = tokenize([batch_dict], cfg)
    emb = embed_and_interleave([batch_dict], tb, w, cfg)
    rows, picked = generate(w, emb, torch.from_numpy(tb.token_masks), cfg, max_length, 0, cfg.text_tokens - 1)
    return torch.stack(rows), picked


def predict_control(w: Dict[str, torch.Tensor], inp: dict, cfg: GatoConfig, action_tokens: int, discrete_n: Optional[int] = None):
    """Continuous (discrete_n None): returns decoded actions (2*bin/n_bins - 1, input_tokenizers.py:32-42);
    discrete: the action index."""
    tb = tokenize([inp], cfg)
    emb = embed_and_interleave([inp], tb, w, cfg)[:, :-action_tokens]
    mask = torch.from_numpy(tb.token_masks)[:, :-action_tokens]
    cont0 = cfg.text_tokens
    disc0 = cfg.text_tokens + cfg.continuous_tokens
    if discrete_n is not None:
        _, picked = generate(w, emb, mask, cfg, 1, disc0, disc0 + discrete_n - 1)
        return picked[0] - disc0
    _, picked = generate(w, emb, mask, cfg, action_tokens, cont0, cont0 + cfg.continuous_tokens - 1)
    t = torch.tensor(picked, dtype=torch.float32) - cont0
    return (2 * t) / cfg.continuous_tokens - 1


# --------------------------------------------------------------------------------------
# deterministic weights (numpy MT19937 => identical on every machine, no torch RNG involved)
# --------------------------------------------------------------------------------------
def weight_shapes(cfg: GatoConfig) -> Dict[str, tuple]:
    """state_dict parameter names/shapes of the reference policy."""
    d, C, p = cfg.embed_dim, cfg.resid_mid_channels, cfg.patch_size
    s: Dict[str, tuple] = {"separator_token": (d,), "transformer.wte.weight": (cfg.wte_rows, d)}
    for i in range(cfg.layers):
        q = f"transformer.h.{i}."
        s.update({
            q + "ln_1.weight": (d,), q + "ln_1.bias": (d,),
            q + "attn.c_attn.weight": (d, 3 * d), q + "attn.c_attn.bias": (3 * d,),
            q + "attn.c_proj.weight": (d, d), q + "attn.c_proj.bias": (d,),
            q + "ln_2.weight": (d,), q + "ln_2.bias": (d,),
            q + "mlp.c_fc.weight": (d, 4 * d), q + "mlp.c_fc.bias": (4 * d,),
            q + "mlp.c_proj.weight": (4 * d, d), q + "mlp.c_proj.bias": (d,),
        })
        if cfg.activation_fn == "geglu":
            s.update({q + "mlp.gated_layer.weight": (4 * d, d), q + "mlp.gated_layer.bias": (4 * d,)})
    s.update({
        "transformer.ln_f.weight": (d,), "transformer.ln_f.bias": (d,),
        "embed_token.weight": (cfg.vocab_size, d),
        "predict_token.weight": (cfg.vocab_size, d),
        "image_embedding.patch_embedding.conv1.weight": (C, 3, 3, 3),
        "image_embedding.patch_embedding.conv1.bias": (C,),
        "image_embedding.patch_embedding.gn2.weight": (C,),
        "image_embedding.patch_embedding.gn2.bias": (C,),
        "image_embedding.patch_embedding.conv2.weight": (3, C, 3, 3),
        "image_embedding.patch_embedding.conv2.bias": (3,),
        "image_embedding.post_embedding_projection.weight": (d, 3 * p * p),
        "image_embedding.post_embedding_projection.bias": (d,),
        "image_embedding.patch_pos_encoding.height_pos_embedding.weight": (cfg.position_vocab_size, d),
        "image_embedding.patch_pos_encoding.width_pos_embedding.weight": (cfg.position_vocab_size, d),
        "pos_embed_observation.weight": (cfg.context_len, d),
    })
    return s


def make_weights(cfg: GatoConfig, seed: int = 0, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's init *distributions* (trajectory_gpt2.py:375-386 for
    the decoder, torch defaults elsewhere) drawn from numpy's MT19937 so fixtures need only a
    seed.  ``perturb`` also randomises LN/GN affine, biases and the separator (zeros/ones at
    init in the reference) so parity tests exercise them."""
    rs = np.random.RandomState(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in weight_shapes(cfg).items():
        if name.startswith("transformer.") and ("ln_" in name):
            if name.endswith("weight"):
                a = 1.0 + (0.1 * rs.standard_normal(shape) if perturb else 0.0)
            else:
                a = 0.1 * rs.standard_normal(shape) if perturb else np.zeros(shape)
        elif name.startswith("transformer."):
            if name.endswith("bias"):
                a = 0.02 * rs.standard_normal(shape) if perturb else np.zeros(shape)
            else:
                a = 0.02 * rs.standard_normal(shape)
        elif name == "separator_token":
            a = 0.5 * rs.standard_normal(shape) if perturb else np.zeros(shape)
        elif "gn2" in name:
            if name.endswith("weight"):
                a = 1.0 + (0.1 * rs.standard_normal(shape) if perturb else 0.0)
            else:
                a = 0.1 * rs.standard_normal(shape) if perturb else np.zeros(shape)
        elif name.endswith("embedding.weight") or name in ("embed_token.weight", "pos_embed_observation.weight"):
            a = rs.standard_normal(shape)  # nn.Embedding default N(0,1)
        else:  # nn.Linear / nn.Conv2d default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            if name.endswith("bias"):
                wshape = weight_shapes(cfg)[name[:-4] + "weight"]
                fan_in = int(np.prod(wshape[1:]))
            else:
                fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            a = rs.uniform(-bound, bound, size=shape)
        out[name] = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), shape)).copy())
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs of the BASELINE.json configs, numpy-seeded
# --------------------------------------------------------------------------------------
def synth_batch(name: str, seed: int = 1234, batch: Optional[int] = None, text_vocab: int = 50257) -> List[dict]:
    """Synthetic dict batches with the shapes/dtypes the reference's tasks emit
    (control_task.py:298-324, text_task.py:47-54, caption_task.py:114-117, vqa_task.py:92-96)."""
    rs = np.random.RandomState(seed)

    def control(T, n_obs, n_act):
        obs = (rs.standard_normal((T, n_obs)) * 3).astype(np.float32)
        act = np.clip(rs.standard_normal((T, n_act)), -1, 1).astype(np.float32)
        return {"continuous_obs": torch.from_numpy(obs), "continuous_actions": torch.from_numpy(act)}

    def atari(T, hw=96):
        img = rs.randint(0, 256, size=(T, 3, hw, hw)).astype(np.float32)
        act = rs.randint(0, 4, size=(T, 1)).astype(np.int32)
        return {"images": torch.from_numpy(img), "discrete_actions": torch.from_numpy(act)}

    def text(n):
        return {"text": rs.randint(0, text_vocab, size=(n,)).tolist()}

    def caption(n_text):
        img = rs.randint(0, 256, size=(1, 3, 224, 224)).astype(np.uint8)
        return {"images": torch.from_numpy(img), "text": torch.from_numpy(rs.randint(0, text_vocab, size=(n_text,)).astype(np.int64))}

    if name == "cfg1":  # HalfCheetah-shaped, B=4, k=240
        return [control(10, 17, 6) for _ in range(batch or 4)]
    if name == "cfg2":  # 3 MuJoCo shapes, B=32, k=240
        shapes = [(17, 6), (11, 3), (17, 6)]
        return [control(240 // (o + a + 1), o, a) for o, a in (shapes[i % 3] for i in range(batch or 32))]
    if name == "cfg3":  # Breakout-shaped, B=32, k=512
        return [atari(512 // 38) for _ in range(batch or 32)]
    if name == "cfg4":  # text, B=16, 1023 ids + separator
        return [text(1023) for _ in range(batch or 16)]
    if name == "cfg5":  # mixed, B=32, k=1024
        B = batch or 32
        q = B // 4
        out = [text(1023) for _ in range(q)]
        out += [caption(32) for _ in range(q)]
        out += [caption(24) for _ in range(q)]
        rest = B - 3 * q
        out += [control(1024 // 24, 17, 6) for _ in range(rest // 2)]
        out += [atari(1024 // 38) for _ in range(rest - rest // 2)]
        return out
    raise ValueError(name)


CONFIGS = {
    "cfg1": dict(embed_dim=128, layers=3, heads=1, context_len=240),
    "cfg2": dict(embed_dim=768, layers=6, heads=24, context_len=240),
    "cfg3": dict(embed_dim=768, layers=6, heads=24, context_len=512),
    "cfg4": dict(embed_dim=768, layers=6, heads=24, context_len=1024),
    "cfg5": dict(embed_dim=768, layers=6, heads=24, context_len=1024),
}
