"""End-to-end optimisation steps on the GPU: Trainer + FusedAdamW (clip + AdamW over the flat arena, LR schedule) against
torch.optim.AdamW driven by the CPU oracle's gradients on the same batches (train.py:127-136, trainer.py:176-188)."""
import copy

import pytest
import torch

from oracle import gato_oracle as O

pytestmark = pytest.mark.gpu


class _Tok:
    def __init__(self, n):
        self.vocab_size = n


class _Replay:
    """Task stub that replays a fixed list of batches (kind 'control')."""
    kind = "control"

    def __init__(self, batches):
        self.batches, self.i = batches, 0

    def sample_batch(self, n, max_tokens=None):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return copy.deepcopy(b)


def _batches(cfg, steps, seed=0, images_at=()):
    """3 control samples per step; steps listed in `images_at` swap the last one for an image-control sample, so the image
    stack has a gradient on those steps only."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for s in range(steps):
        b = [dict(continuous_obs=torch.randn(6, 5, generator=g) * 3,
                  continuous_actions=torch.randn(6, 2, generator=g).clamp(-1, 1)) for _ in range(3)]
        if s in images_at:
            b[-1] = dict(images=torch.randint(0, 256, (3, 3, 32, 32), generator=g).float(),
                         discrete_actions=torch.randint(0, 4, (3, 1), generator=g).to(torch.int32))
        out.append(b)
    return out


@pytest.mark.parametrize("lean,images_at", [(True, ()), (False, ()), (True, (1, 3))])
def test_trainer_steps_match_torch_adamw_on_oracle_gradients(lean, images_at):
    """Against UNMODIFIED torch.optim.AdamW under zero_grad(set_to_none=True): parameters without a gradient this step
    (transformer.wte always; the image stack on image-free steps) are skipped -- no decay, no moment update, own step count."""
    from neko_b200.policy import GatoPolicy
    from neko_b200.training.arguments import TrainingArgs
    from neko_b200.training.trainer import FusedAdamW, Trainer, lr_at_step
    cfg = O.GatoConfig(embed_dim=64, layers=2, heads=2, context_len=64, text_tokens=120)
    w = O.make_weights(cfg, seed=11)
    m = GatoPolicy(device="cuda", embed_dim=64, layers=2, heads=2, dropout=0.0, resid_mid_channels=128, context_len=64,
                   text_tokenizer=_Tok(120))
    m.transformer.drop.p = 0.0
    m.load_state_dict(w, strict=False)
    m.materialize_logits = not lean
    steps = 4
    args = TrainingArgs()
    args.batch_size, args.sequence_length = 3, 64
    args.learning_rate, args.init_lr, args.min_factor = 1e-3, 1e-4, 10.0
    args.warmup_steps, args.training_steps = 2, 8
    args.grad_norm_clip, args.disable_grad_clip = 1.0, False
    args.gradient_accumulation_steps = 1
    args.text_prop = args.caption_prop = args.vqa_prop = 0.0
    batches = _batches(cfg, steps, images_at=images_at)
    m.eval()        # eval-mode patch positions (deterministic) on both sides; dropout is 0 anyway

    class _OneSample(_Replay):
        """Trainer draws control samples one at a time (trainer.py:211-247): hand out the batch sample by sample."""
        def __init__(self, batches):
            super().__init__([s for b in batches for s in b])

        def sample_batch(self, n, max_tokens=None):
            s = self.batches[self.i % len(self.batches)]
            self.i += 1
            return [copy.deepcopy(s)]

    opt = FusedAdamW(m, lr=args.learning_rate, betas=(args.beta_1, args.beta_2), eps=args.adam_eps, weight_decay=args.weight_decay)
    tr = Trainer(m, opt, [_OneSample(batches)], args)
    losses = []
    for _ in range(steps):          # Trainer.train() would switch to train mode
        losses.append(tr.train_step()[0])

    # reference: oracle forward/backward on CPU + torch AdamW with the same schedule and clipping
    for t in w.values():
        t.requires_grad_(True)
    used = [t for n, t in w.items() if n != "transformer.wte.weight" and not n.startswith("image_embedding.")]
    ropt = torch.optim.AdamW(list(w.values()), lr=args.learning_rate, betas=(args.beta_1, args.beta_2), eps=args.adam_eps,
                             weight_decay=args.weight_decay)
    rlosses = []
    for s in range(steps):
        lr = lr_at_step(s, warmup_steps=args.warmup_steps, training_steps=args.training_steps, base_lr=args.learning_rate,
                        init_lr=args.init_lr, min_lr=args.learning_rate / args.min_factor, cosine_decay=True)
        for gI in ropt.param_groups:
            gI["lr"] = lr
        out = O.forward(w, batches[s], cfg, compute_loss=True, training=False)
        ropt.zero_grad(set_to_none=True)
        out.loss.backward()
        torch.nn.utils.clip_grad_norm_([t for t in w.values() if t.grad is not None], args.grad_norm_clip)
        ropt.step()
        rlosses.append(float(out.loss.detach()))
    assert len(used) > 10
    for a, b in zip(losses, rlosses):
        assert abs(a - b) <= 2e-3 * abs(b), (losses, rlosses)
    if not images_at:       # same kind of batch every step: the loss must go down
        assert losses[-1] < losses[0]
    # parameters after 4 AdamW steps: updates are O(lr) per step whatever the gradient scale, so compare in units of lr
    sd = m.state_dict()
    worst = 0.0
    w0 = O.make_weights(cfg, seed=11)
    for n, t in w.items():
        diff = (sd[n].detach().cpu() - t.detach()).abs()
        worst = max(worst, float(diff.mean()) / args.learning_rate)
        if n == "transformer.wte.weight" or (n.startswith("image_embedding.") and not images_at):
            assert torch.equal(sd[n].detach().cpu(), w0[n]), f"{n} has no gradient and must not move (no weight decay)"
    assert worst < 0.5, worst
    if images_at:
        assert opt.steps["image_embedding.patch_embedding.conv1.weight"] == len(images_at) and opt.steps["embed_token.weight"] == steps
        assert "transformer.wte.weight" not in opt.steps


def test_checkpoint_resume_is_exact(tmp_path):
    """Model + optimiser state round trip: 2 steps, save, reload into a fresh model / optimiser, 2 more steps
    == 4 uninterrupted steps (up to the run-to-run noise of atomically accumulated gradients)."""
    from neko_b200.policy import GatoPolicy
    from neko_b200.training.arguments import TrainingArgs
    from neko_b200.training.trainer import FusedAdamW, Trainer, load_checkpoint, save_checkpoint
    cfg = O.GatoConfig(embed_dim=64, layers=2, heads=2, context_len=64, text_tokens=120)
    w = O.make_weights(cfg, seed=11)
    args = TrainingArgs()
    args.batch_size, args.sequence_length = 3, 64
    args.learning_rate, args.init_lr, args.min_factor = 1e-3, 1e-4, 10.0
    args.warmup_steps, args.training_steps = 2, 8
    args.gradient_accumulation_steps = 1
    args.text_prop = args.caption_prop = args.vqa_prop = 0.0
    batches = _batches(cfg, 4)
    flat = [s for b in batches for s in b]

    class _Seq(_Replay):
        def __init__(self, samples, start=0):
            super().__init__(samples)
            self.i = start

        def sample_batch(self, n, max_tokens=None):
            s = self.batches[self.i % len(self.batches)]
            self.i += 1
            return [copy.deepcopy(s)]

    def fresh():
        m = GatoPolicy(device="cuda", embed_dim=64, layers=2, heads=2, dropout=0.0, resid_mid_channels=128, context_len=64,
                       text_tokenizer=_Tok(120))
        m.transformer.drop.p = 0.0
        m.load_state_dict(w, strict=False)
        m.materialize_logits = False
        opt = FusedAdamW(m, lr=args.learning_rate, betas=(args.beta_1, args.beta_2), eps=args.adam_eps, weight_decay=args.weight_decay)
        return m, opt

    m1, o1 = fresh()
    t1 = Trainer(m1, o1, [_Seq(flat)], args)
    t1.prefetch = False
    l_full = [l for l, _ in t1.train(4)]

    m2, o2 = fresh()
    t2 = Trainer(m2, o2, [_Seq(flat)], args)
    t2.prefetch = False
    l_a = [l for l, _ in t2.train(2)]
    path = save_checkpoint(m2, str(tmp_path), "ckpt", args, optimizer=o2, step=t2.steps)
    m3, o3 = fresh()
    step = load_checkpoint(m3, path, optimizer=o3)
    assert step == 2 and o3.t == 2
    t3 = Trainer(m3, o3, [_Seq(flat, start=6)], args)     # 2 steps x 3 samples already consumed
    t3.prefetch = False
    t3.steps = step
    l_b = [l for l, _ in t3.train(2)]
    # atomically accumulated gradients are not bit-reproducible run to run: compare at the level of that noise
    for a, b in zip(l_a + l_b, l_full):
        assert abs(a - b) <= 1e-4 * abs(b), (l_a + l_b, l_full)
    assert (m3._param_arena - m1._param_arena).abs().mean().item() <= 0.02 * args.learning_rate
    assert (o3.exp_avg - o1.exp_avg).abs().max().item() <= 1e-3 * o1.exp_avg.abs().max().item()
    assert (o3.exp_avg_sq - o1.exp_avg_sq).abs().max().item() <= 1e-3 * o1.exp_avg_sq.abs().max().item()
    # and a resume WITHOUT the optimiser state is visibly different (the moments matter)
    m4, o4 = fresh()
    load_checkpoint(m4, path, optimizer=None)
    t4 = Trainer(m4, o4, [_Seq(flat, start=6)], args)
    t4.prefetch = False
    t4.steps = 2
    t4.train(2)
    assert (m4._param_arena - m1._param_arena).abs().mean().item() > 0.05 * args.learning_rate
