"""Parameter containers and host-side integer logic of the image front end, mirroring
gato/policy/embeddings.py (ImageEmbedding :7-61, PatchPosEncoding :63-110, ResidualBlock_V2 :111-131).

The modules own the parameters under the reference's state_dict names; the arithmetic runs in
csrc/patch_embed.cu + the tcgen05 GEMM (see GatoPolicy._image_forward).  The patch-position BINS are
computed here with the very torch calls the reference makes, which keeps the train-mode global-CPU-RNG
stream bit-identical.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class _ConvParams(nn.Module):
    def __init__(self, out_ch: int, in_ch: int, k: int = 3):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_ch, in_ch, k, k))
        self.bias = nn.Parameter(torch.empty(out_ch))
        # nn.Conv2d default init (kaiming_uniform a=sqrt(5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
        bound = 1.0 / math.sqrt(in_ch * k * k)
        nn.init.uniform_(self.weight, -bound, bound)
        nn.init.uniform_(self.bias, -bound, bound)


class _AffineParams(nn.Module):
    def __init__(self, n: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))


class _LinearParams(nn.Module):
    def __init__(self, in_f: int, out_f: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f))
        bound = 1.0 / math.sqrt(in_f)
        nn.init.uniform_(self.weight, -bound, bound)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_f))
            nn.init.uniform_(self.bias, -bound, bound)
        else:
            self.register_parameter("bias", None)


class _EmbeddingParams(nn.Module):
    def __init__(self, n: int, d: int, std: float = 1.0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n, d))
        nn.init.normal_(self.weight, 0.0, std)

    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        """Row lookup for callers outside the fused path (the predict_* loops embed one generated token at a time,
        gato_policy.py:465,600); training gathers rows inside tokenize_embed_kernel instead."""
        return self.weight.detach()[idx.to(self.weight.device).long()]


class ResidualBlock_V2(nn.Module):
    """embeddings.py:111-131: conv1 (3->C) / gn2 (groups, C) / conv2 (C->3); gn1 is Identity."""

    def __init__(self, mid_channels: int = 128, num_groups: int = 32):
        super().__init__()
        if mid_channels % num_groups != 0:
            # nn.GroupNorm raises the same way in the reference (default 132 fails)
            raise ValueError("num_channels must be divisible by num_groups")
        self.mid_channels = mid_channels
        self.num_groups = num_groups
        self.conv1 = _ConvParams(mid_channels, 3)
        self.gn2 = _AffineParams(mid_channels)
        self.conv2 = _ConvParams(3, mid_channels)


class PatchPosEncoding(nn.Module):
    def __init__(self, position_vocab_size: int = 128, embed_dim: int = 768):
        super().__init__()
        self.position_vocab_size = position_vocab_size
        self.embed_dim = embed_dim
        self.height_pos_embedding = _EmbeddingParams(position_vocab_size, embed_dim)
        self.width_pos_embedding = _EmbeddingParams(position_vocab_size, embed_dim)

    def positions(self, n_height: int, n_width: int):
        """Row / column bins, embeddings.py:80-100 verbatim in behaviour (CPU tensors)."""
        h_linspace = torch.linspace(0, 1, n_height + 1)
        w_linspace = torch.linspace(0, 1, n_width + 1)
        h_intervals = torch.stack([h_linspace[:-1], h_linspace[1:]]).T
        w_intervals = torch.stack([w_linspace[:-1], w_linspace[1:]]).T
        h_intervals = (h_intervals * self.position_vocab_size).to(dtype=torch.int32)
        w_intervals = (w_intervals * self.position_vocab_size).to(dtype=torch.int32)
        if self.training:
            # one draw per row, then per column, on the global CPU generator
            h_pos = torch.tensor([torch.randint(low=int(iv[0]), high=int(iv[1]), size=()) for iv in h_intervals])
            w_pos = torch.tensor([torch.randint(low=int(iv[0]), high=int(iv[1]), size=()) for iv in w_intervals])
        else:
            h_intervals[:, 1] = h_intervals[:, 1] - 1
            w_intervals[:, 1] = w_intervals[:, 1] - 1
            h_pos = h_intervals.mean(dim=-1, dtype=torch.float32).round().to(dtype=torch.int32)
            w_pos = w_intervals.mean(dim=-1, dtype=torch.float32).round().to(dtype=torch.int32)
        return h_pos.to(torch.int32), w_pos.to(torch.int32)


class ImageEmbedding(nn.Module):
    def __init__(self, embed_dim=768, patch_size=16, resid_mid_channels=128, num_groups=32,
                 position_vocab_size=128, use_pos_encoding=True):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.patch_embedding = ResidualBlock_V2(mid_channels=resid_mid_channels, num_groups=num_groups)
        self.post_embedding_projection = _LinearParams(patch_size * patch_size * 3, embed_dim)
        self.use_pos_encoding = use_pos_encoding
        self.patch_pos_encoding = PatchPosEncoding(position_vocab_size=position_vocab_size, embed_dim=embed_dim)

    def forward(self, x, normalize=True):
        """[T,3,H,W] -> [T, n_h*n_w, embed_dim] (embeddings.py:28-61); routed through the owning policy."""
        owner = getattr(self, "_owner", None)
        if owner is None:
            raise RuntimeError("ImageEmbedding must be owned by a GatoPolicy (it runs on the policy's CUDA engine)")
        if not normalize:
            raise NotImplementedError("normalize=False is not used on the reference's hot path")
        return owner()._embed_images_standalone(x)
