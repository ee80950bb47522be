"""ContinuousTokenizer with the reference's interface (gato/policy/input_tokenizers.py:9-42).

``encode`` runs the bit-exact CUDA discretiser (csrc/tokenize.cu) -- the same code the fused
tokenise/embed kernel uses; ``decode`` is host-side scalar arithmetic used only at inference
(gato_policy.py:612)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import SampleDesc, TokParams, _p, check, load, stream_ptr


class ContinuousTokenizer:
    def __init__(self, use_mu_law=True, mu=100, M=256, n_bins=1024, offset=None):
        self.use_mu_law = use_mu_law
        self.mu = mu
        self.M = M
        self.n_bins = n_bins
        self.offset = offset

    def encode(self, tensor: torch.Tensor) -> torch.Tensor:
        """fp32 tensor (CUDA) -> int32 token ids of the same shape (input_tokenizers.py:17-30)."""
        if not tensor.is_cuda:
            raise _lib.NekoError("ContinuousTokenizer.encode needs a CUDA tensor (no CPU path in neko_b200)")
        x = tensor.detach().to(torch.float32).contiguous().reshape(-1)
        n = x.numel()
        # one "sample" with one timestep of n continuous values, no embeddings
        desc = SampleDesc()
        desc.n_timesteps = 1
        if self.use_mu_law:
            desc.n_cobs = n
        else:
            desc.n_cact = n
        # the observation block precedes the separator, the action block follows it
        width = n + 1
        descs = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(x.device)
        prm = TokParams(mu=float(self.mu), M=float(self.M), n_bins=int(self.n_bins),
                        cont_start=int(self.offset or 0), disc_start=0, vocab=1 << 30, use_pos=0,
                        seq_len=width, width=width, ctx_rows=0)
        tokens = torch.empty(width, dtype=torch.int64, device=x.device)
        tm = torch.empty(width, dtype=torch.float32, device=x.device)
        mk = torch.empty(width, dtype=torch.float32, device=x.device)
        check(load().neko_tokenize_embed_fwd(_p(descs), C.c_int(1), C.c_int(4), C.byref(prm), _p(x), _p(None), _p(None),
                                             _p(None), _p(None), _p(None), _p(tokens), _p(tm), _p(mk), _p(None), _p(None),
                                             stream_ptr()), "neko_tokenize_embed_fwd")
        ids = tokens[:n] if self.use_mu_law else tokens[1:]
        return ids.to(torch.int32).reshape(tensor.shape)

    def decode(self, tensor):
        if self.use_mu_law:
            raise Exception("mu-law encoding only expected with values which are not predicted")
        if self.offset is not None:
            tensor -= self.offset
        return (2 * tensor) / self.n_bins - 1
