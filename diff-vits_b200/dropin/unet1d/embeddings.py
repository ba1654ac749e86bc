"""Embedding modules of the 1-D UNet (reference: unet1d/embeddings.py).

`TextTimeEmbedding` / `AttentionPooling` only depend on the speech prompt, so on the denoising path they are
evaluated once per utterance (PyTorch ops, hoisted out of the step loop) and their result is handed to the CUDA
engine as `aug`. `model3.py:40` also imports `TextTimeEmbedding` for the VITS prior, so it has to be a fully
functional nn.Module. `Timesteps` + `TimestepEmbedding` hold parameters; their math runs in the engine's
time-embedding kernel (csrc/elementwise.cu), the PyTorch forwards below are kept for API completeness (VITS uses).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .activations import get_activation


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool = False,
                           downscale_freq_shift: float = 1, scale: float = 1, max_period: int = 10000):
    """Sinusoidal embedding of (possibly fractional) timesteps, [N] -> [N, embedding_dim].
    Reference: unet1d/embeddings.py:24-64 (frequencies exp(-ln(max_period) * j / (half - shift)), sin | cos,
    optionally swapped to cos | sin, zero pad for odd dims)."""
    if timesteps.dim() != 1:
        raise AssertionError("Timesteps should be a 1d-array")
    half = embedding_dim // 2
    freqs = time_frequencies(half, downscale_freq_shift, max_period, timesteps.device)
    arg = scale * (timesteps[:, None].float() * freqs[None, :])
    parts = (torch.cos(arg), torch.sin(arg)) if flip_sin_to_cos else (torch.sin(arg), torch.cos(arg))
    emb = torch.cat(parts, dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def time_frequencies(half_dim: int, downscale_freq_shift: float = 0, max_period: int = 10000, device=None):
    """The fp32 frequency table, computed with the same fp32 op sequence as the reference so the CUDA kernel can
    consume bit-identical values (embeddings.py:42-47)."""
    exponent = -math.log(max_period) * torch.arange(start=0, end=half_dim, dtype=torch.float32, device=device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    return torch.exp(exponent)


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    """Linear -> act -> Linear (reference: embeddings.py:157-201)."""

    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu", out_dim: int = None,
                 post_act_fn: Optional[str] = None, cond_proj_dim=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.cond_proj = nn.Linear(cond_proj_dim, in_channels, bias=False) if cond_proj_dim is not None else None
        self.act = get_activation(act_fn)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)
        self.post_act = get_activation(post_act_fn) if post_act_fn is not None else None

    def forward(self, sample, condition=None):
        if condition is not None:
            sample = sample + self.cond_proj(condition)
        sample = self.linear_2(self.act(self.linear_1(sample)))
        return self.post_act(sample) if self.post_act is not None else sample


class AttentionPooling(nn.Module):
    """One learned-position class token (mean of the sequence + positional embedding) attends over
    [token, sequence] with `num_heads` heads; returns the attended token (reference: embeddings.py:499-546).
    Padded (zeroed) prompt rows take part exactly as in the reference."""

    def __init__(self, num_heads, embed_dim, dtype=None):
        super().__init__()
        self.dtype = dtype
        self.positional_embedding = nn.Parameter(torch.randn(1, embed_dim) / embed_dim**0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim, dtype=self.dtype)
        self.q_proj = nn.Linear(embed_dim, embed_dim, dtype=self.dtype)
        self.v_proj = nn.Linear(embed_dim, embed_dim, dtype=self.dtype)
        self.num_heads = num_heads
        self.dim_per_head = embed_dim // self.num_heads

    def forward(self, x):
        bs, length, width = x.shape
        h, d = self.num_heads, self.dim_per_head
        token = x.mean(dim=1, keepdim=True) + self.positional_embedding.to(x.dtype)
        seq = torch.cat([token, x], dim=1)                                   # [bs, 1+length, width]
        s = 1.0 / math.sqrt(math.sqrt(d))
        q = self.q_proj(token).view(bs, 1, h, d).permute(0, 2, 1, 3) * s     # [bs, h, 1, d]
        k = self.k_proj(seq).view(bs, 1 + length, h, d).permute(0, 2, 1, 3) * s
        v = self.v_proj(seq).view(bs, 1 + length, h, d).permute(0, 2, 1, 3)
        w = torch.softmax((q @ k.transpose(-1, -2)).float(), dim=-1).type(q.dtype)   # [bs, h, 1, 1+length]
        return (w @ v).reshape(bs, width)                                    # heads concatenated: [bs, width]


class TextTimeEmbedding(nn.Module):
    """LayerNorm -> AttentionPooling -> Linear -> LayerNorm (reference: embeddings.py:421-434)."""

    def __init__(self, encoder_dim: int, time_embed_dim: int, num_heads: int = 64):
        super().__init__()
        self.norm1 = nn.LayerNorm(encoder_dim)
        self.pool = AttentionPooling(num_heads, encoder_dim)
        self.proj = nn.Linear(encoder_dim, time_embed_dim)
        self.norm2 = nn.LayerNorm(time_embed_dim)

    def forward(self, hidden_states):
        return self.norm2(self.proj(self.pool(self.norm1(hidden_states))))
