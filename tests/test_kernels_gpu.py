"""Kernel-level parity: every C-ABI entry point against a plain torch fp32 restatement of the same op,
on the same seeded inputs.  GPU only (-m gpu)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from neko_b200 import _lib, ops as _ops
    _lib.require_device()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


# ---------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    (128, 128, 64), (128, 128, 256), (256, 384, 768), (200, 72, 104), (1000, 200, 72), (130, 520, 1032),
    (7680, 2304, 768), (7680, 768, 3072), (1024, 768, 768),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_layouts(ops, M, N, K, a_mn, b_mn):
    # leading dimensions must be multiples of 8 elements: pad the storage, slice the logical view
    def pad8(n):
        return (n + 7) // 8 * 8
    a_log = _rand((M, K), 1, 0.5)
    b_log = _rand((N, K), 2, 0.5)
    if a_mn:
        a_store = torch.zeros(K, pad8(M), device="cuda", dtype=torch.bfloat16)
        a_store[:, :M] = a_log.t().to(torch.bfloat16)
        a = a_store[:, :M]
    else:
        a_store = torch.zeros(M, pad8(K), device="cuda", dtype=torch.bfloat16)
        a_store[:, :K] = a_log.to(torch.bfloat16)
        a = a_store[:, :K]
    if b_mn:
        b_store = torch.zeros(K, pad8(N), device="cuda", dtype=torch.bfloat16)
        b_store[:, :N] = b_log.t().to(torch.bfloat16)
        b = b_store[:, :N]
    else:
        b_store = torch.zeros(N, pad8(K), device="cuda", dtype=torch.bfloat16)
        b_store[:, :K] = b_log.to(torch.bfloat16)
        b = b_store[:, :K]
    ref = a_log.to(torch.bfloat16).float() @ b_log.to(torch.bfloat16).float().t()
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=ops.EPI_F32, M=M, N=N, K=K)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    tol = 2e-3 * math.sqrt(K / 64) + 1e-4
    assert err < tol, f"max err {err} (tol {tol})"


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("M,N,K,b_mn", [(7680, 768, 3072, False), (1000, 2304, 768, False), (640, 768, 768, True)])
def test_gemm_tile_width_192(ops, monkeypatch, pair, M, N, K, b_mn):
    """The 192-column tile (single CTA and CTA pair), forced through NEKO_GEMM_BN / NEKO_GEMM_PAIR."""
    monkeypatch.setenv("NEKO_GEMM_BN", "192")
    monkeypatch.setenv("NEKO_GEMM_PAIR", str(pair))
    a = _rand((M, K), 61, 1.0, torch.bfloat16)
    b = _rand((K, N) if b_mn else (N, K), 62, 0.05, torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, b_mn=b_mn, epilogue=ops.EPI_BF16, out=out)
    ref = a.float() @ (b.float() if b_mn else b.float().t())
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("M,N,K", [(768, 768, 7680), (768, 3072, 7680), (768, 2304, 4000), (3072, 768, 15808), (256, 128, 2048)])
def test_gemm_split_k_weight_gradients(ops, M, N, K):
    """wgrad shape: few output tiles, K = the token axis -> split-K with fp32 RED reduction, both overwrite and
    accumulate semantics."""
    x = _rand((K, M), 71, 0.5, torch.bfloat16)   # activations [tokens, in]
    dy = _rand((K, N), 72, 0.5, torch.bfloat16)  # gradients  [tokens, out]
    ref = x.float().t() @ dy.float()
    out = torch.full((M, N), 7.0, device="cuda")
    ops.gemm(x, dy, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=out)
    tol = 2e-3 * math.sqrt(K / 64) + 1e-4
    assert (out - ref).abs().max().item() < tol
    ops.gemm(x, dy, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=out, accumulate=True)
    assert (out - 2 * ref).abs().max().item() < 2 * tol


def test_gemm_epilogues(ops):
    M, N, K = 384, 256, 192
    a = _rand((M, K), 3, 0.5, torch.bfloat16)
    b = _rand((N, K), 4, 0.5, torch.bfloat16)
    bias = _rand((N,), 5)
    resid = _rand((M, N), 6)
    pre = _rand((M, N), 7, 1.0, torch.bfloat16)
    acc = a.float() @ b.float().t()

    out = ops.gemm(a, b, epilogue=ops.EPI_BF16, bias=bias)
    assert (out.float() - (acc + bias)).abs().max().item() < 0.08
    out = ops.gemm(a, b, epilogue=ops.EPI_BF16)
    assert (out.float() - acc).abs().max().item() < 0.08

    out = ops.gemm(a, b, epilogue=ops.EPI_F32, bias=bias)
    assert (out - (acc + bias)).abs().max().item() < 2e-3
    base = resid.clone()
    out = ops.gemm(a, b, epilogue=ops.EPI_F32, out=base, accumulate=True)
    assert (out - (acc + resid)).abs().max().item() < 2e-3

    pre_o, act_o = ops.gemm(a, b, epilogue=ops.EPI_GELU_BF16, bias=bias)
    assert (pre_o.float() - (acc + bias)).abs().max().item() < 0.08
    assert (act_o.float() - torch.nn.functional.gelu(acc + bias)).abs().max().item() < 0.08

    x = resid.clone()
    out = ops.gemm(a, b, epilogue=ops.EPI_RESID_F32, bias=bias, aux=x, out=x)  # in place on the residual stream
    assert (out - (acc + bias + resid)).abs().max().item() < 2e-3

    x = resid.clone()
    out, out_bf = ops.gemm(a, b, epilogue=ops.EPI_RESID_F32_BF16, bias=bias, aux=x, out=x)
    assert (out - (acc + bias + resid)).abs().max().item() < 2e-3
    assert (out_bf.float() - out).abs().max().item() < 0.08

    p = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(p).sum().backward()
    out = ops.gemm(a, b, epilogue=ops.EPI_DGELU_BF16, aux=pre)
    assert (out.float() - acc * p.grad).abs().max().item() < 0.1


@pytest.mark.parametrize("M,N,K,pair", [(7680, 768, 3072, 1), (7680, 768, 768, 0), (7680, 768, 2304, 1), (5000, 768, 1024, 0)])
def test_gemm_stream_k(ops, monkeypatch, M, N, K, pair):
    """Stream-K scheduling (csrc/gemm.cu WorkIter): a tile's k-range shared by two workers through the workspace.  The decoder's
    N = 768 launches at the cfg2 token count (90 pair tiles on 74 pairs) with every epilogue that takes part, against fp32 torch and
    against the classic schedule; twice in a row (the counters re-arm themselves)."""
    monkeypatch.setenv("NEKO_GEMM_PAIR", str(pair))
    a = _rand((M, K), 3, 0.5, torch.bfloat16)
    b = _rand((N, K), 4, 0.5, torch.bfloat16)
    bias = _rand((N,), 5)
    resid = _rand((M, N), 6)
    pre = _rand((M, N), 7, 1.0, torch.bfloat16)
    ref = a.float() @ b.float().t()
    tol = 2e-3 * (K / 192) ** 0.5
    outs = {}
    for sk in ("0", "1", "1"):
        monkeypatch.setenv("NEKO_GEMM_STREAMK", sk)
        o16 = ops.gemm(a, b, epilogue=ops.EPI_BF16, bias=bias)
        x = resid.clone()
        o32 = ops.gemm(a, b, epilogue=ops.EPI_RESID_F32, bias=bias, aux=x, out=x)
        od = ops.gemm(a, b, epilogue=ops.EPI_DGELU_BF16, aux=pre)
        torch.cuda.synchronize()
        assert (o32 - (ref + bias + resid)).abs().max().item() < tol, sk
        assert (o16.float() - (ref + bias)).abs().max().item() < 0.01 * (ref + bias).abs().max().item() + 0.05, sk
        outs.setdefault(sk, []).append((o16.clone(), o32.clone(), od.clone()))
    c0 = outs["0"][0]
    for c1 in outs["1"]:
        assert (c0[1] - c1[1]).abs().max().item() < tol            # different summation order, same result within fp32 rounding
        assert (c0[0].float() - c1[0].float()).abs().max().item() <= 0.02 * c0[0].float().abs().max().item()
        assert (c0[2].float() - c1[2].float()).abs().max().item() <= 0.02 * c0[2].float().abs().max().item() + 1e-3
    assert torch.equal(outs["1"][0][1], outs["1"][1][1])        # deterministic: partials are added in a fixed order


def test_gemm_gelu_tanh_epilogues(ops):
    """gelu_new (tanh form, pretrained GPT-2: gato_policy.py:79-95) in the GELU / GELU' epilogues.  The two GELU forms differ by
    < 5e-4, below one 16-bit ulp, so the check is statistical: with B = I the pre-activation is exact, rounding errors average
    out over 64k elements and the mean signed error tells the forms apart."""
    M, N = 512, 128
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(M, N, generator=g) * 0.8 + 1.8).cuda().to(torch.bfloat16)      # where tanh - erf (2.4e-4) and its derivative (6.6e-4) have one sign
    eye = torch.eye(N, device="cuda", dtype=torch.bfloat16)
    F = torch.nn.functional
    xf = x.float()
    gap = (F.gelu(xf, approximate="tanh") - F.gelu(xf)).mean().item()
    assert abs(gap) > 1e-4
    for tanh in (True, False):
        pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        act = torch.empty(M, N, device="cuda", dtype=torch.float16)
        ops.gemm(x, eye, epilogue=ops.EPI_GELU_BF16, out=pre, out2=act, gelu_tanh=tanh)
        assert torch.equal(pre, x)
        ref = F.gelu(xf, approximate="tanh") if tanh else F.gelu(xf)
        other = F.gelu(xf) if tanh else F.gelu(xf, approximate="tanh")
        assert (act.float() - ref).abs().max().item() < 2.5e-3
        assert abs((act.float() - ref).mean().item()) < 0.3 * abs(gap) < abs((act.float() - other).mean().item())
    # derivative: acc = 1 (ones @ I / N trick: A = ones, B = I), aux = x
    ones = torch.ones(M, N, device="cuda", dtype=torch.bfloat16)
    for tanh in (True, False):
        p = xf.clone().requires_grad_(True)
        (F.gelu(p, approximate="tanh") if tanh else F.gelu(p)).sum().backward()
        q = xf.clone().requires_grad_(True)
        (F.gelu(q) if tanh else F.gelu(q, approximate="tanh")).sum().backward()
        dgap = (p.grad - q.grad).mean().item()
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        ops.gemm(ones, eye, epilogue=ops.EPI_DGELU_BF16, aux=x, out=out, gelu_tanh=tanh)
        assert (out.float() - p.grad).abs().max().item() < 2e-3
        assert abs((out.float() - p.grad).mean().item()) < 0.3 * abs(dgap) < abs((out.float() - q.grad).mean().item())


@pytest.mark.parametrize("a_dt,b_dt", [(torch.float16, torch.float16)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (False, True)])
def test_gemm_fp16_operands(ops, a_dt, b_dt, a_mn, b_mn):
    """Forward runs fp16 x fp16 (mixed f16/bf16 operands trap on tcgen05 kind::f16 and are rejected by the API)."""
    M, N, K = 384, 512, 320
    a_log = _rand((M, K), 61, 0.5).to(a_dt)
    b_log = _rand((N, K), 62, 0.5).to(b_dt)
    a = a_log.t().contiguous() if a_mn else a_log
    b = b_log.t().contiguous() if b_mn else b_log
    ref = a_log.float() @ b_log.float().t()
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=ops.EPI_F32)
    assert (out - ref).abs().max().item() < 3e-3
    out16 = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=ops.EPI_BF16, out_dtype=torch.float16)
    assert out16.dtype == torch.float16 and (out16.float() - ref).abs().max().item() < 0.02
    bias = _rand((N,), 63)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    act16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    actbf = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epilogue=ops.EPI_GELU_BF16, out=pre, out2=act16, out3=actbf, bias=bias)
    gref = torch.nn.functional.gelu(ref + bias)
    assert (act16.float() - gref).abs().max().item() < 0.02 and (actbf.float() - gref).abs().max().item() < 0.08
    with pytest.raises(Exception):
        ops.gemm(a, b.to(torch.bfloat16), a_mn=a_mn, b_mn=b_mn)


def test_gemm_lm_head_shape(ops):
    """K=768, N = 52305 rows of W padded to a 52352-column output (OOB rows of B zero-filled by TMA)."""
    M, K, V, Vp = 256, 768, 52305, 52352
    h = _rand((M, K), 8, 1.0, torch.bfloat16)
    w_store = torch.zeros(Vp, K, device="cuda", dtype=torch.bfloat16)  # operands must cover the padded extent
    w_store[:V] = _rand((V, K), 9, 0.02, torch.bfloat16)
    w = w_store[:V]
    out = torch.full((M, Vp), float("nan"), device="cuda")
    ops.gemm(h, w_store, epilogue=ops.EPI_F32, out=out, N=Vp)
    ref = h.float() @ w.float().t()
    assert (out[:, :V] - ref).abs().max().item() < 5e-3
    assert (out[:, V:] == 0).all()
    # dgrad through the head: dh = dlogits[M,Vp] @ W[V,K] with K-tail on the vocabulary
    dl = torch.zeros(M, Vp, device="cuda", dtype=torch.bfloat16)
    dl[:, :V] = _rand((M, V), 10, 0.01, torch.bfloat16)
    dh = ops.gemm(dl, w, b_mn=True, epilogue=ops.EPI_F32, K=V, N=K)
    ref = dl[:, :V].float() @ w.float()
    assert (dh - ref).abs().max().item() < 5e-3
    # wgrad: dW[V,K] = dlogits^T h
    dw = ops.gemm(dl, h, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, M=V, N=K, K=M)
    ref = dl[:, :V].float().t() @ h.float()
    assert (dw - ref).abs().max().item() < 5e-3


# ---------------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,d", [(37, 32), (1000, 64), (513, 128), (2048, 768), (100, 1024), (64, 2048)])
def test_layernorm(ops, N, d):
    x = _rand((N, d), 11, 2.0) + 0.5
    g = _rand((d,), 12, 0.2) + 1.0
    b = _rand((d,), 13, 0.2)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5)
    y2 = torch.empty(N, d, device="cuda", dtype=torch.bfloat16)
    y16, _, _ = ops.layernorm_fwd(x, g, b, 1e-5, out_dtype=torch.float16, y2=y2)
    assert torch.equal(y2, y)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5)
    assert (y.float() - ref).abs().max().item() < 0.05
    assert y16.dtype == torch.float16 and (y16.float() - ref).abs().max().item() < 0.01
    assert (mean - x.mean(-1)).abs().max().item() < 1e-5
    dy = _rand((N, d), 14, 1.0, torch.bfloat16)
    ref.backward(dy.float())
    resid = _rand((N, d), 15)
    dx = resid.clone()
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    dx_bf = torch.empty(N, d, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(d, device="cuda")
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, dx_bf, dx_colsum=cs)
    assert (dx - (resid + xr.grad)).abs().max().item() < 2e-4
    assert (cs - dx.sum(0)).abs().max().item() < 1e-3 * math.sqrt(N) + 1e-3
    assert (dx_bf.float() - dx).abs().max().item() < 0.05
    assert (dg - gr.grad).abs().max().item() < 2e-3 * math.sqrt(N)
    assert (db - br.grad).abs().max().item() < 2e-3 * math.sqrt(N)


# ---------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------
def _attn_ref(qkv, first_valid, H, S_valid):
    """Reference semantics (trajectory_gpt2.py:163-188 + 663-679) in fp32 with -1e4 fills."""
    B, S, three_d = qkv.shape
    d = three_d // 3
    dh = d // H
    q, k, v = qkv.float().split(d, dim=2)
    q = q.reshape(B, S, H, dh).permute(0, 2, 1, 3)
    k = k.reshape(B, S, H, dh).permute(0, 2, 3, 1)
    v = v.reshape(B, S, H, dh).permute(0, 2, 1, 3)
    pos = torch.arange(S, device=qkv.device)
    mask = ((pos[None, :] >= first_valid[:, None]) & (pos[None, :] < S_valid)).float()
    w = torch.matmul(q, k) / (float(dh) ** 0.5)
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool, device=qkv.device))[None, None]
    w = torch.where(causal, w, torch.tensor(-1e4, device=qkv.device))
    w = w + ((1.0 - mask) * -10000.0)[:, None, None, :]
    p = torch.softmax(w, dim=-1)
    return torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B, S, d), mask


@pytest.mark.parametrize("B,S,H,dh,S_valid", [(2, 64, 2, 32, 64), (3, 100, 3, 32, 100), (2, 240, 24, 32, 240), (2, 130, 1, 128, 130),
                                               (2, 96, 4, 16, 80), (2, 200, 2, 64, 200), (1, 494, 24, 32, 494)])
@pytest.mark.parametrize("tc", [0, 1])
def test_attention_fwd_bwd(ops, monkeypatch, tc, B, S, H, dh, S_valid):
    monkeypatch.setenv("NEKO_ATTN_TC", str(tc))   # 1: tcgen05 forward where the shape allows it (dh 32 with even H, dh 64)
    d = H * dh
    qkv = _rand((B, S, 3 * d), 21, 1.0, torch.bfloat16)
    fv = torch.tensor([0, 17, 70][:B], dtype=torch.int32).clamp(max=S_valid - 1).cuda()
    out, lse = ops.attention_fwd(qkv, fv, H, S_valid)
    x = qkv.float().requires_grad_(True)
    ref, mask = _attn_ref(x, fv, H, S_valid)
    live = mask.bool()[:, :, None].expand(B, S, d)
    err = ((out.float() - ref) * live).abs().max().item()
    assert err < 0.03, err
    assert (out.float() * (~live)).abs().max().item() == 0.0  # padded rows are zeros
    dout = _rand((B, S, d), 22, 1.0, torch.bfloat16) * live
    (ref * dout.float()).sum().backward()
    dqkv = ops.attention_bwd(qkv, out, dout, lse, fv, H, S_valid)
    o2 = torch.empty_like(out)
    out16, lse16 = ops.attention_fwd(qkv, fv, H, S_valid, out_dtype=torch.float16, out2=o2)
    assert ((out16.float() - ref) * live).abs().max().item() < 0.03 and torch.equal(lse16, lse) and torch.equal(o2, out)
    dqkv16 = ops.attention_bwd(qkv, out16, dout, lse, fv, H, S_valid)
    assert (dqkv16.float() - dqkv.float()).abs().max().item() < 0.03 * max(x.grad.abs().max().item(), 1.0)
    gref = x.grad
    scale = gref.abs().max().item()
    err = (dqkv.float() - gref).abs().max().item()
    assert err < 0.03 * max(scale, 1.0), (err, scale)


# ---------------------------------------------------------------------------------------------------
# masked cross entropy
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,ld", [(2208, 2208), (52305, 52352), (1001, 1001)])
def test_masked_ce(ops, V, ld):
    B, S = 3, 40
    N = B * S
    store = torch.zeros(N, ld, device="cuda")
    store[:, :V] = _rand((N, V), 31, 2.0)
    tokens = torch.randint(0, V, (B, S), generator=torch.Generator().manual_seed(5)).cuda()
    g = torch.Generator().manual_seed(6)
    sel = (torch.rand(B, S - 1, generator=g) < 0.3)
    rows = (torch.arange(B)[:, None] * S + torch.arange(S - 1)[None, :])[sel].to(torch.int32).cuda()
    loss, row_lse, row_loss = ops.masked_ce_fwd(store, V, rows, tokens.reshape(-1))
    z = store[:, :V].clone().requires_grad_(True)
    tgt = tokens.reshape(-1)[rows.long() + 1]
    ref = torch.nn.functional.cross_entropy(z[rows.long()], tgt)
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    (ref * 0.5).backward()
    gs = torch.tensor(0.5, device="cuda")
    dl = torch.zeros(N, ld, device="cuda", dtype=torch.bfloat16)
    ops.masked_ce_bwd(store, V, rows, tokens.reshape(-1), row_lse, gs, dl)
    assert (dl[:, :V].float() - z.grad).abs().max().item() < 2 ** -8 * z.grad.abs().max().item() + 1e-6  # bf16 rounding
    dlc = torch.zeros(rows.numel(), ld, device="cuda", dtype=torch.bfloat16)
    ops.masked_ce_bwd(store, V, rows, tokens.reshape(-1), row_lse, gs, dlc, flags=ops.CE_DLOGITS_COMPACT)
    assert torch.equal(dlc, dl[rows.long()])


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def test_cast_colsum_rows(ops):
    x = _rand((1000, 777), 41)
    xb = ops.cast_bf16(x.contiguous())
    assert torch.equal(xb, x.to(torch.bfloat16))
    y = _rand((3000, 776), 42, 1.0, torch.bfloat16)
    out = torch.zeros(776, device="cuda")
    ops.colsum(y, out)
    assert (out - y.float().sum(0)).abs().max().item() < 0.05
    ops.colsum(y, out, accumulate=True)
    assert (out - 2 * y.float().sum(0)).abs().max().item() < 0.1
    rows = torch.randperm(3000)[:500].to(torch.int32).cuda()
    g = torch.empty(500, 776, device="cuda", dtype=torch.bfloat16)
    ops.gather_rows(y, rows, 776, g)
    assert torch.equal(g, y[rows.long()])
    dst = _rand((3000, 776), 43)
    ref = dst.clone()
    ref[rows.long()] += g.float()
    ops.scatter_rows_add(g, rows, 776, dst)
    assert (dst - ref).abs().max().item() < 1e-6


def test_adamw_clip(ops):
    n = 100_003
    p = _rand((n,), 51)
    g = _rand((n,), 52, 3.0)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    n_cast = 60_001
    w16 = torch.zeros(n_cast, device="cuda", dtype=torch.float16)
    wbf = torch.zeros(n_cast, device="cuda", dtype=torch.bfloat16)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    for step in range(1, 4):
        pr.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        ss = torch.zeros((), device="cuda")
        ops.sumsq(g, ss)
        ops.adamw_step(p, g, m, v, 1e-3, 0.9, 0.95, 1e-8, 0.1, step, ss, 1.0, w_f16=w16, w_bf16=wbf, n_cast=n_cast)
    assert (p - pr.detach()).abs().max().item() < 1e-5
    # fused refresh of the 16-bit operand copies of the first n_cast parameters
    assert torch.equal(w16, p[:n_cast].to(torch.float16)) and torch.equal(wbf, p[:n_cast].to(torch.bfloat16))


# ---------------------------------------------------------------------------------------------------
# image patch ResNet block (embeddings.py:28-61,111-131)
# ---------------------------------------------------------------------------------------------------
def _patch_block_torch(img, w1, b1, gw, gb, w2, b2, groups):
    import torch.nn.functional as F
    n, _, H, W = img.shape
    x = (img.float() / 255.0 * 2 - 1) / 4.0
    x = x.view(n, 3, H // 16, 16, W // 16, 16).permute(0, 2, 4, 1, 3, 5).reshape(-1, 3, 16, 16)
    h = F.conv2d(F.gelu(x), w1, b1, padding=1)
    h = F.gelu(F.group_norm(h, groups, gw, gb, eps=1e-5))
    return (x + F.conv2d(h, w2, b2, padding=1)).reshape(x.shape[0], -1)


@pytest.mark.parametrize("groups", [32, 16, 8, 64])
@pytest.mark.parametrize("u8", [True, False])
def test_patch_resblock(ops, groups, u8):
    import ctypes as C
    from neko_b200 import _lib
    from neko_b200.ops import _p, stream_ptr
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(7 + groups)
    n, H, W, Cm = 5, 48, 64, 128
    img = torch.randint(0, 256, (n, 3, H, W), generator=g, dtype=torch.uint8)
    img = (img if u8 else img.float() + 0.25).cuda()
    w1 = _rand((Cm, 3, 3, 3), 1, 0.3); b1 = _rand((Cm,), 2, 0.2)
    gw = 1.0 + _rand((Cm,), 3, 0.2); gb = _rand((Cm,), 4, 0.2)
    w2 = _rand((3, Cm, 3, 3), 5, 0.05); b2 = _rand((3,), 6, 0.1)
    P = n * (H // 16) * (W // 16)
    out16 = torch.empty(P, 768, dtype=torch.float16, device="cuda")
    outbf = torch.empty(P, 768, dtype=torch.bfloat16, device="cuda")
    stats = torch.empty(P, groups, 2, device="cuda")
    _lib.check(lib.neko_patch_resblock_fwd(_p(img), C.c_int(int(u8)), C.c_int(n), C.c_int(H), C.c_int(W), C.c_int(16), C.c_int(Cm),
                                           C.c_int(groups), _p(w1), _p(b1), _p(gw), _p(gb), _p(w2), _p(b2), _p(out16), _p(outbf),
                                           _p(stats), stream_ptr()), "fwd")
    params = [t.clone().requires_grad_(True) for t in (w1, b1, gw, gb, w2, b2)]
    ref = _patch_block_torch(img, *params, groups)
    err = (out16.float() - ref).abs().max().item()
    assert err < 6e-3, err                      # bf16 conv operands, fp32 accumulation; outputs are O(0.3)
    assert ((outbf.float() - ref).abs() - ref.abs() * 2.0 ** -8).max().item() < 6e-3
    # GroupNorm statistics of the conv1 output
    with torch.no_grad():
        import torch.nn.functional as F
        x = (img.float() / 255.0 * 2 - 1) / 4.0
        x = x.view(n, 3, H // 16, 16, W // 16, 16).permute(0, 2, 4, 1, 3, 5).reshape(-1, 3, 16, 16)
        hh = F.conv2d(F.gelu(x), w1, b1, padding=1).view(P, groups, -1)
        assert (stats[..., 0] - hh.mean(-1)).abs().max().item() < 2e-3
        assert ((stats[..., 1] - (hh.var(-1, unbiased=False) + 1e-5).rsqrt()) / stats[..., 1]).abs().max().item() < 2e-2
    # backward
    dy = _rand((P, 768), 9, 1.0, torch.bfloat16)
    ref.backward(dy.float())
    grads = [torch.zeros_like(t) for t in (w1, b1, gw, gb, w2, b2)]
    _lib.check(lib.neko_patch_resblock_bwd(_p(img), C.c_int(int(u8)), C.c_int(n), C.c_int(H), C.c_int(W), C.c_int(16), C.c_int(Cm),
                                           C.c_int(groups), _p(w1), _p(b1), _p(gw), _p(gb), _p(w2), _p(stats), _p(dy),
                                           _p(grads[0]), _p(grads[1]), _p(grads[2]), _p(grads[3]), _p(grads[4]), _p(grads[5]),
                                           stream_ptr()), "bwd")
    for name, got, p in zip(("w1", "b1", "gw", "gb", "w2", "b2"), grads, params):
        want = p.grad
        rel = ((got - want).norm() / want.norm()).item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        assert rel < 2e-2 and cos > 0.9995, (name, rel, cos)
    # gradients accumulate into their destination
    _lib.check(lib.neko_patch_resblock_bwd(_p(img), C.c_int(int(u8)), C.c_int(n), C.c_int(H), C.c_int(W), C.c_int(16), C.c_int(Cm),
                                           C.c_int(groups), _p(w1), _p(b1), _p(gw), _p(gb), _p(w2), _p(stats), _p(dy),
                                           _p(grads[0]), _p(grads[1]), _p(grads[2]), _p(grads[3]), _p(grads[4]), _p(grads[5]),
                                           stream_ptr()), "bwd")
    assert ((grads[0] - 2 * params[0].grad).norm() / params[0].grad.norm()).item() < 4e-2


# ---------------------------------------------------------------------------------------------------
# dropout primitives
# ---------------------------------------------------------------------------------------------------
def _seed(a, b):
    return torch.tensor([a, b], dtype=torch.int32, device="cuda")


def test_dropout_mask_apply_and_statistics(ops):
    from neko_b200._lib import Dropout
    s = _seed(123, 456)
    for p in (0.1, 0.5):
        dr = Dropout.make(s, 7, p)
        keep = ops.dropout_mask(1000, 777, dr, "cuda")
        rate = keep.float().mean().item()
        assert abs(rate - (1 - p)) < 5e-3, rate
        # columns / rows are not correlated with each other
        assert abs(keep[:, ::2].float().mean().item() - keep[:, 1::2].float().mean().item()) < 1e-2
        assert (keep[:-1] == keep[1:]).float().mean().item() < (p * p + (1 - p) * (1 - p)) + 1e-2
        x = _rand((1000, 777), 5)
        y = x.clone()
        ops.dropout_apply(y, dr)
        assert torch.equal(y, x * keep.float() * dr.scale)
        # another stream / another seed -> another mask; same (seed, stream) -> same mask
        assert not torch.equal(keep, ops.dropout_mask(1000, 777, Dropout.make(s, 8, p), "cuda"))
        assert not torch.equal(keep, ops.dropout_mask(1000, 777, Dropout.make(_seed(124, 456), 7, p), "cuda"))
        assert torch.equal(keep, ops.dropout_mask(1000, 777, Dropout.make(s, 7, p), "cuda"))
    assert Dropout.make(s, 1, 0.0).thr16 == 0


def test_gemm_residual_dropout_and_layernorm_bwd_branch_mask(ops):
    from neko_b200._lib import Dropout
    M, N, K = 520, 768, 256
    dr = Dropout.make(_seed(9, 10), 6, 0.25)
    a = _rand((M, K), 1, 1.0, torch.float16)
    b = _rand((K, N), 2, 0.05, torch.float16)
    bias = _rand((N,), 3)
    res = _rand((M, N), 4)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, b_mn=True, epilogue=ops.EPI_RESID_F32, out=out, aux=res, bias=bias, drop=dr)
    mult = ops.dropout_mask(M, N, dr, "cuda").float() * dr.scale
    ref = res + (a.float() @ b.float() + bias) * mult
    assert (out - ref).abs().max().item() < 2e-3
    # LN backward: dx_bf16 / dx_colsum carry mask * scale * dx, the fp32 residual gradient stays unmasked
    d = N
    x = _rand((M, d), 5)
    gamma = 1 + _rand((d,), 6, 0.1)
    dy = _rand((M, d), 7, 1.0, torch.bfloat16)
    dx0 = _rand((M, d), 8)
    mean = x.mean(1)
    rstd = (x.var(1, unbiased=False) + 1e-5).rsqrt()
    dxa, dxb_ = dx0.clone(), dx0.clone()
    dga, dba, dgb, dbb = (torch.zeros(d, device="cuda") for _ in range(4))
    c_a, c_b = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    b_a = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    b_b = torch.empty_like(b_a)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dxa, dga, dba, b_a, dx_colsum=c_a)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dxb_, dgb, dbb, b_b, dx_colsum=c_b, branch_drop=dr)
    assert torch.equal(dxa, dxb_) and (dga - dgb).abs().max().item() <= 1e-4 * float(dga.abs().max())   # dgamma: atomics
    assert torch.equal(b_b, (dxb_ * mult).to(torch.bfloat16))
    assert (c_b - (dxb_ * mult).sum(0)).abs().max().item() < 2e-2 * max(1.0, float((dxb_ * mult).sum(0).abs().max()))


@pytest.mark.parametrize("B,S,H,dh,S_valid", [(2, 200, 3, 32, 200), (2, 130, 2, 64, 120), (1, 64, 1, 128, 64), (3, 96, 2, 16, 96),
                                               (2, 300, 4, 32, 300)])
@pytest.mark.parametrize("tc", [0, 1])
def test_attention_dropout_fwd_bwd(ops, monkeypatch, tc, B, S, H, dh, S_valid):
    from neko_b200._lib import Dropout
    monkeypatch.setenv("NEKO_ATTN_TC", str(tc))
    d = H * dh
    dr = Dropout.make(_seed(31, 32), 5, 0.2)
    qkv = _rand((B, S, 3 * d), 11, 0.7, torch.bfloat16)
    fv = torch.tensor([0, 17, 5][:B], dtype=torch.int32, device="cuda")
    out, lse = ops.attention_fwd(qkv, fv, H, S_valid, drop=dr)
    mult = (ops.dropout_mask(B * H * S, S, dr, "cuda").float() * dr.scale).view(B, H, S, S)
    qf = qkv.float().requires_grad_(True)
    q, k, v = (t.view(B, S, H, dh).transpose(1, 2) for t in qf.split(d, dim=2))
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    idx = torch.arange(S, device="cuda")
    ok = (idx[None, :] <= idx[:, None])[None, None] & (idx[None, None, None, :] >= fv[:, None, None, None])
    sc = sc.masked_fill(~ok, float("-inf"))
    pr = torch.softmax(sc, -1).nan_to_num(0.0) * mult
    ref = (pr @ v).transpose(1, 2).reshape(B, S, d)
    live = (idx[None, :] >= fv[:, None]) & (idx[None, :] < S_valid)
    assert (out.float() - ref)[live].abs().max().item() < 3e-2
    dout = _rand((B, S, d), 12, 1.0, torch.bfloat16) * live[..., None]
    (ref * live[..., None]).backward(dout.float())
    dqkv = ops.attention_bwd(qkv, out, dout, lse, fv, H, S_valid, drop=dr)
    g = qf.grad
    rel = ((dqkv.float() - g)[live].norm() / g[live].norm()).item()
    assert rel < 3e-2, rel


@pytest.mark.parametrize("B,S,H,dh,S_valid,fvs", [(3, 512, 4, 32, 512, (0, 130, 300)), (2, 300, 2, 64, 260, (5, 129)),
                                                   (3, 1024, 2, 32, 1024, (0, 255, 1000)), (2, 128, 2, 32, 128, (0, 127))])
@pytest.mark.parametrize("tc", [1, 0])
def test_attention_tensor_core_forward_long(ops, monkeypatch, tc, B, S, H, dh, S_valid, fvs):
    """Both forward paths -- tcgen05/TMEM/TMA (attention_tc.cu, NEKO_ATTN_TC=1) and mma.sync (default) -- on several
    128-key tiles, left padding beyond the first tile, right padding."""
    monkeypatch.setenv("NEKO_ATTN_TC", str(tc))
    d = H * dh
    qkv = _rand((B, S, 3 * d), 23, 1.0, torch.bfloat16)
    fv = torch.tensor(list(fvs), dtype=torch.int32).cuda()
    o2 = torch.empty(B, S, d, device="cuda", dtype=torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, fv, H, S_valid, out_dtype=torch.float16, out2=o2)
    ref, mask = _attn_ref(qkv.float(), fv, H, S_valid)
    live = mask.bool()[:, :, None].expand(B, S, d)
    assert ((out.float() - ref) * live).abs().max().item() < 0.03
    assert ((o2.float() - ref) * live).abs().max().item() < 0.03
    assert (out.float() * (~live)).abs().max().item() == 0.0
    # log-sum-exp of the live rows against the fp32 scores
    q, k, _ = (t.view(B, S, H, dh).transpose(1, 2) for t in qkv.float().split(d, dim=2))
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    idx = torch.arange(S, device="cuda")
    ok = (idx[None, :] <= idx[:, None])[None, None] & (idx[None, None, None, :] >= fv[:, None, None, None])
    ref_lse = torch.logsumexp(sc.masked_fill(~ok, float("-inf")), dim=-1)
    lrow = mask.bool()[:, None, :].expand(B, H, S)
    assert (lse - ref_lse)[lrow].abs().max().item() < 2e-2
    assert torch.isinf(lse[~lrow]).all()
    # and the mma.sync backward consumes it
    dout = _rand((B, S, d), 24, 1.0, torch.bfloat16) * live
    dqkv = ops.attention_bwd(qkv, out, dout, lse, fv, H, S_valid)
    assert torch.isfinite(dqkv.float()).all()


@pytest.mark.parametrize("V,ld", [(2208, 2208), (52305, 52352), (1001, 1004)])
def test_masked_ce_fused_matches_two_pass(ops, V, ld):
    """ce_fused_kernel (loss + unscaled dlogits in one pass) against the separate forward / backward kernels, and the
    conditional scale in backward."""
    B, S = 3, 40
    N = B * S
    g = torch.Generator(device="cpu").manual_seed(V)
    store = torch.zeros(N, ld, device="cuda")
    store[:, :V] = (torch.randn(N, V, generator=g) * 3).cuda()
    tokens = torch.randint(0, V, (N,), generator=g).cuda()
    rows = torch.arange(0, N - 1, 3, dtype=torch.int32).cuda()
    n = rows.numel()
    ldd = (V + 63) // 64 * 64
    loss_ref, lse_ref, _ = ops.masked_ce_fwd(store, V, rows, tokens)
    one = torch.ones((), device="cuda")
    d_ref = torch.full((n, ldd), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.masked_ce_bwd(store, V, rows, tokens, lse_ref, one, d_ref, flags=ops.CE_DLOGITS_COMPACT | ops.CE_ZERO_PAD)
    d_new = torch.full((n, ldd), 7.0, device="cuda", dtype=torch.bfloat16)
    out = ops.masked_ce_fused(store, V, rows, tokens, d_new, flags=ops.CE_DLOGITS_COMPACT | ops.CE_ZERO_PAD)
    assert out is not None
    loss, lse = out
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    assert (lse - lse_ref).abs().max().item() <= 1e-4
    assert (d_new.float() - d_ref.float()).abs().max().item() <= 2.0 ** -8 * d_ref.float().abs().max().item() + 1e-8
    assert float(d_new[:, V:].abs().max()) == 0.0 if ldd > V else True
    # upstream scalar: exactly 1 leaves the buffer untouched, anything else scales it
    keep = d_new.clone()
    ops.ce_scale_grad(d_new, one)
    assert torch.equal(d_new, keep)
    ops.ce_scale_grad(d_new, torch.full((), 0.25, device="cuda"))
    assert torch.equal(d_new, (keep.float() * 0.25).to(torch.bfloat16))
