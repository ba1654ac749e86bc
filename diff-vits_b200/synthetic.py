"""Seeded synthetic workload of the benchmark and the parity tests: UNet constructor
arguments of the diffusion encoder (model3.py:887-896 with config.json:40-46), inputs seed 0 N(0,1) fp32, linear
betas 1e-4..2e-2 x 1000. Product-side module: the oracle re-exports these so both sides see identical inputs."""
from __future__ import annotations

import torch

UNET_KW = dict(in_channels=228, out_channels=100, block_out_channels=(128, 256, 384, 512), norm_num_groups=8,
               cross_attention_dim=128, attention_head_dim=8, addition_embed_type="text",
               resnet_time_scale_shift="scale_shift")


def betas_linear(n: int = 1000) -> torch.Tensor:
    """model3.py:935-942 linear_beta_schedule computed in fp64 then stored fp32 (:978-992)."""
    scale = 1000 / n
    return torch.linspace(scale * 0.0001, scale * 0.02, n, dtype=torch.float64).to(torch.float32)


def make_inputs(B: int, T: int, Tp: int, seed: int = 0, x_ch: int = 100, cond_ch: int = 128, ehs_dim: int = 128):
    g = torch.Generator().manual_seed(seed)
    x_T = torch.randn(B, x_ch, T, generator=g)
    cond = torch.randn(B, cond_ch, T, generator=g)
    ehs = torch.randn(B, Tp, ehs_dim, generator=g)
    return x_T, cond, ehs


def ragged_mask(B: int, Tp: int, step: int = 100):
    """lengths [Tp, Tp-step, Tp-2*step, ...] (>= 1) and the matching keep-mask; masked ehs rows must be zeroed by the
    caller the way Diffusion_Encoder does (model3.py:905-906)."""
    lens = torch.tensor([max(1, Tp - i * step) for i in range(B)], dtype=torch.int32)
    mask = torch.arange(Tp)[None, :] < lens[:, None]
    return lens, mask


def make_prompt(B: int, Tp: int, seed: int = 0, n_mels: int = 100) -> torch.Tensor:
    """Speech-prompt mel [B, 100, Tp] ~ N(0,1): the PromptEncoder's input (model3.py:906), for the arms that run it."""
    g = torch.Generator().manual_seed(seed + 7919)
    return torch.randn(B, n_mels, Tp, generator=g)
