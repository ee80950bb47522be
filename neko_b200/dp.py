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
    def __init__(self, arena: torch.Tensor, group=None, bucket_bytes: int = 64 << 20, average: bool = True,
                 mode: str = "overlap", compress: str = "none", backend: str = "auto"):
        # mode "overlap": buckets fire from inside backward on a side stream; "tail": everything is reduced after backward
        # (backward then replays from its CUDA graph, and no communication kernel competes with the persistent GEMMs)
        # compress "bf16": gradients cross NVLink as bf16 (half the bytes), summed by NCCL in bf16, unpacked with the
        # 1/world factor in fp32
        assert mode in ("overlap", "tail") and compress in ("none", "bf16")
        self.mode, self.compress = mode, compress
        self._pack = None
        # backend "p2p": own all-reduce kernel over IPC-mapped peer memory (csrc/p2p_allreduce.cu), co-resident with the
        # GEMMs of backward; "nccl": torch.distributed all_reduce (also the gloo path of the CPU tests)
        if backend == "auto":
            backend = "p2p" if (arena.is_cuda and dist.is_initialized() and 2 <= dist.get_world_size(group) <= 8
                                and compress == "none" and average) else "nccl"
        self.backend = backend
        self._p2p = None
        # the p2p kernels keep all their bookkeeping on the device and may be captured into the backward CUDA graph
        self.capturable = backend == "p2p"
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
        if self.mode == "tail":
            self._pending.append((lo, hi))
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
                if self.backend == "p2p":
                    if self._p2p is None:
                        self._p2p = _P2PState(self.arena, self.group)
                    self._p2p.all_reduce(lo, hi, 1.0 / self.world if self.average else 1.0)
                elif self.compress == "bf16":
                    if self._pack is None or self._pack.numel() < self.arena.numel():
                        self._pack = torch.empty(self.arena.numel(), dtype=torch.bfloat16, device=self.arena.device)
                    pk = self._pack[lo:hi]
                    pk.copy_(view)
                    dist.all_reduce(pk, op=op, group=self.group)
                    view.copy_(pk)
                    if self.average:
                        view.mul_(1.0 / self.world)
                else:
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
            if self.mode == "tail" and self._pending:   # one all-reduce per contiguous live span
                spans = sorted(self._pending)
                merged = [list(spans[0])]
                for lo, hi in spans[1:]:
                    if lo <= merged[-1][1]:
                        merged[-1][1] = max(merged[-1][1], hi)
                    else:
                        merged.append([lo, hi])
                self._pending = []
                for lo, hi in merged:
                    for a, b in self._minus_skipped(lo, hi):
                        self._launch(a, b)
            else:
                self._flush()
        if self._cuda:
            torch.cuda.current_stream(self.arena.device).wait_stream(self._stream)

    def begin_step(self):
        self.launched = []

    def finish_tail(self):
        """mode "tail": called by the policy after backward (eager or graph replay) -- reduce everything now."""
        if self.world == 1 or not self.enabled:
            return
        self.launched = []
        self._pending = list(self._tail_ranges)
        self.finish()

    @contextlib.contextmanager
    def no_sync(self):
        prev = self.enabled
        self.enabled = False
        try:
            yield
        finally:
            self.enabled = prev
            self._pending = []


class _P2PState:
    """Peer mappings + barrier bookkeeping of csrc/p2p_allreduce.cu for one arena (include/neko_b200.h: neko_ipc_*,
    neko_p2p_allreduce_f32).  Built collectively: every rank of the group must construct it at the same point."""

    def __init__(self, arena: torch.Tensor, group=None):
        import ctypes as C
        from ._lib import check, load
        self._C, self._check, self._lib = C, check, load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.arena = arena
        dev = arena.device
        self.sig = torch.zeros(64, dtype=torch.int32, device=dev)          # [world] signal words (+ padding)
        self.state = torch.zeros(2, dtype=torch.int32, device=dev)         # launch counter, CTA arrival counter
        self.n_ctas = int(__import__('os').environ.get('NEKO_P2P_CTAS', 0)) or int(self._lib.neko_sm_count())
        torch.cuda.synchronize(dev)

        def export(t):
            h = (C.c_ubyte * 64)()
            off = C.c_longlong(0)
            check(self._lib.neko_ipc_export(C.c_void_p(t.data_ptr()), h, C.byref(off)), "neko_ipc_export")
            return bytes(h), int(off.value)

        mine = {"arena": export(arena), "sig": export(self.sig), "numel": arena.numel(), "pid": __import__("os").getpid()}
        every = [None] * self.world
        dist.all_gather_object(every, mine, group=group)
        if any(e["numel"] != arena.numel() for e in every):
            raise RuntimeError("p2p all-reduce: the ranks' gradient arenas differ in size")
        self._opened = []

        def imp(pair):
            h = (C.c_ubyte * 64).from_buffer_copy(pair[0])
            out = C.c_void_p(0)
            check(self._lib.neko_ipc_import(h, C.c_longlong(pair[1]), C.byref(out)), "neko_ipc_import")
            self._opened.append((out.value, pair[1]))
            return out.value

        bufs, sigs = [], []
        for r, e in enumerate(every):
            if r == self.rank:
                bufs.append(arena.data_ptr())
                sigs.append(self.sig.data_ptr())
            else:
                bufs.append(imp(e["arena"]))
                sigs.append(imp(e["sig"]))
        self.bufs = (C.c_void_p * self.world)(*bufs)
        self.sigs = (C.c_void_p * self.world)(*sigs)
        dist.barrier(group=group)      # every signal buffer is zeroed and mapped before anybody launches

    def all_reduce(self, lo: int, hi: int, scale: float):
        C = self._C
        self._check(self._lib.neko_p2p_allreduce_f32(self.bufs, self.sigs, C.c_void_p(self.state.data_ptr()), C.c_int(self.rank),
                                                     C.c_int(self.world), C.c_longlong(lo), C.c_longlong(hi),
                                                     C.c_float(scale), C.c_int(self.n_ctas),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "neko_p2p_allreduce_f32")


def attach(policy, group=None, bucket_bytes: int = 64 << 20, no_text_tokens: bool = False, mode: str = "overlap",
           compress: str = "none", backend: str = "auto") -> GradSynchronizer:
    """Wire a GradSynchronizer to a GatoPolicy: buckets fire from inside ``loss.backward()``.

    ``no_text_tokens``: the caller guarantees that no rank ever feeds text tokens (a task mix with
    text_prop = caption_prop = vqa_prop = 0, trainer.py:134): rows [0, text_tokens) of ``embed_token`` -- 154 MB of the
    last, non-overlappable bucket -- then have zero gradient everywhere and are left out of the all-reduce, like the
    never-used ``transformer.wte``.  The result is identical to the dense all-reduce."""
    sync = GradSynchronizer(policy._grad_arena, group=group, bucket_bytes=bucket_bytes, mode=mode, compress=compress, backend=backend)
    if sync.backend == "p2p" and sync.world > 1:
        sync._p2p = _P2PState(policy._grad_arena, group)     # collective: map the peers' arenas now, not inside backward
    if no_text_tokens:
        o = policy._offs["embed_token.weight"]
        sync.skip_ranges.append((o, o + policy.text_tokens * policy.embed_dim))
        w = policy._offs["transformer.wte.weight"]
        sync.skip_ranges.append((w, w + policy._params["transformer.wte.weight"].numel()))
    if mode == "tail":
        # nothing fires inside backward: it keeps replaying from its CUDA graph; the policy calls sync.finish_tail() after it
        policy.grad_ready_hook = None
        sync._tail_ranges = _live_ranges(policy)
    else:
        policy.grad_ready_hook = sync.on_range_ready
    policy._grad_sync = sync
    return sync


def _live_ranges(policy):
    """Arena span of every parameter that can carry a gradient (everything but transformer.wte)."""
    total = policy._grad_arena.numel()
    w = policy._offs["transformer.wte.weight"]
    return [(0, w)] + ([(w + 64, total)] if w + 64 < total else [])


def broadcast_parameters(policy, src: int = 0, group=None):
    """All ranks start from rank ``src``'s weights (what DDP does at wrap time)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(policy._param_arena, src=src, group=group)
        policy._bf16_versions = None
