"""Data-parallel gradient synchronisation for the flat gradient arena.

Replaces what the reference gets from Accelerate -> torch DDP (train.py:26-40,107; trainer.py:176-186): a
SUM all-reduce of every gradient followed by a division by the world size, bucketed and overlapped with
backward.  Here the buckets are static arena ranges that become final in reverse execution order
(head -> blocks L-1..0 -> embeddings/image stack); GatoPolicy._engine_backward announces each range as soon
as its last kernel is enqueued, the synchroniser orders a side stream after that point and launches one NCCL
all-reduce per bucket over NVLink/NVSwitch.  There is no unused-parameter bitmap round: unused parameters are
zero ranges of the arena.  Per-rank loss stays a local mean (DDP semantics).

``no_sync`` (gradient accumulation, trainer.py:176) simply skips the launches until the last micro-step.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class GradSynchronizer:
    def __init__(self, arena: torch.Tensor, group=None, bucket_bytes: int = 64 << 20, average: bool = True):
        self.arena = arena
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // arena.element_size())
        self.average = average
        self.enabled = True
        self._pending: List[Tuple[int, int]] = []
        self._works = []
        self._cuda = arena.is_cuda
        self._stream = torch.cuda.Stream(device=arena.device) if self._cuda else None
        self.launched: List[Tuple[int, int]] = []   # bucket log of the last step (tests)
        # arena ranges the caller guarantees to be zero on EVERY rank (e.g. the text rows of embed_token when no task
        # of the mix produces text tokens): averaging zeros is a no-op, so they are left out of the all-reduce
        self.skip_ranges: List[Tuple[int, int]] = []

    # -- hook called by the backward engine -----------------------------------------------------------
    def on_range_ready(self, lo: int, hi: int):
        if not self.enabled or self.world == 1:
            return
        # coalesce small ranges (LN / bias vectors) with their neighbours up to the bucket size
        if self._pending and self._pending[-1][1] == lo and (hi - self._pending[-1][0]) <= self.bucket_elems:
            self._pending[-1] = (self._pending[-1][0], hi)
        else:
            self._flush()
            self._pending.append((lo, hi))
        if self._pending[-1][1] - self._pending[-1][0] >= self.bucket_elems:
            self._flush()

    def _flush(self):
        for lo, hi in self._pending:
            for a, b in self._minus_skipped(lo, hi):
                # split oversized ranges so that the first chunk can start while later ones are still queued
                for s in range(a, b, self.bucket_elems):
                    self._launch(s, min(b, s + self.bucket_elems))
        self._pending = []

    def _minus_skipped(self, lo: int, hi: int):
        out = [(lo, hi)]
        for slo, shi in self.skip_ranges:
            nxt = []
            for a, b in out:
                if shi <= a or slo >= b:
                    nxt.append((a, b))
                else:
                    if a < slo:
                        nxt.append((a, slo))
                    if shi < b:
                        nxt.append((shi, b))
            out = nxt
        return out

    def _launch(self, lo: int, hi: int):
        view = self.arena[lo:hi]
        self.launched.append((lo, hi))
        op = dist.ReduceOp.SUM
        if self._cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.arena.device))
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                dist.all_reduce(view, op=op, group=self.group)
                if self.average:
                    view.mul_(1.0 / self.world)
        else:
            dist.all_reduce(view, op=op, group=self.group)
            if self.average:
                view.mul_(1.0 / self.world)

    # -- end of backward -----------------------------------------------------------------------------------
    def finish(self):
        """Drain pending ranges and make the compute stream wait for the communication stream."""
        if self.world == 1:
            return
        if self.enabled:
            self._flush()
        if self._cuda:
            torch.cuda.current_stream(self.arena.device).wait_stream(self._stream)

    def begin_step(self):
        self.launched = []

    @contextlib.contextmanager
    def no_sync(self):
        prev = self.enabled
        self.enabled = False
        try:
            yield
        finally:
            self.enabled = prev
            self._pending = []


def attach(policy, group=None, bucket_bytes: int = 64 << 20, no_text_tokens: bool = False) -> GradSynchronizer:
    """Wire a GradSynchronizer to a GatoPolicy: buckets fire from inside ``loss.backward()``.

    ``no_text_tokens``: the caller guarantees that no rank ever feeds text tokens (a task mix with
    text_prop = caption_prop = vqa_prop = 0, trainer.py:134): rows [0, text_tokens) of ``embed_token`` -- 154 MB of the
    last, non-overlappable bucket -- then have zero gradient everywhere and are left out of the all-reduce, like the
    never-used ``transformer.wte``.  The result is identical to the dense all-reduce."""
    sync = GradSynchronizer(policy._grad_arena, group=group, bucket_bytes=bucket_bytes)
    if no_text_tokens:
        o = policy._offs["embed_token.weight"]
        sync.skip_ranges.append((o, o + policy.text_tokens * policy.embed_dim))
        w = policy._offs["transformer.wte.weight"]
        sync.skip_ranges.append((w, w + policy._params["transformer.wte.weight"].numel()))
    policy.grad_ready_hook = sync.on_range_ready
    policy._grad_sync = sync
    return sync


def broadcast_parameters(policy, src: int = 0, group=None):
    """All ranks start from rank ``src``'s weights (what DDP does at wrap time)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(policy._param_arena, src=src, group=group)
        policy._bf16_versions = None
