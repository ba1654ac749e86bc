"""Activation lookup (reference: unet1d/activations.py:4-12)."""
from torch import nn


def get_activation(act_fn: str) -> nn.Module:
    if act_fn in ("swish", "silu"):
        return nn.SiLU()
    if act_fn == "mish":
        return nn.Mish()
    if act_fn == "gelu":
        return nn.GELU()
    raise ValueError(f"Unsupported activation function: {act_fn}")
