"""Attention parameter holder (reference: unet1d/attention_processor.py:26-355; processor AttnProcessor2_0 :971-1052).
q/k/v projections carry no bias, the output projection does; softmax(q k^T / sqrt(d) + bias) v runs in the fused
flash-style kernel csrc/attention.cu."""
from typing import Optional

from torch import nn

from ._holder import ParamHolder


class AttnProcessor2_0:
    """Marker kept for API parity: the math it names (SDPA) is what csrc/attention.cu implements."""


class Attention(ParamHolder):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias=False, upcast_attention: bool = False, upcast_softmax: bool = False,
                 out_bias: bool = True, scale_qk: bool = True, **unsupported):
        super().__init__()
        bad = {k: v for k, v in unsupported.items() if v not in (None, False, 1.0, 32, 1e-5)}
        if bad or bias or not out_bias or not scale_qk:
            raise ValueError(f"unsupported Attention options for the Diff-VITS denoiser: {bad}")
        inner_dim = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head**-0.5
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim, bias=True), nn.Dropout(dropout)])
        self.processor = AttnProcessor2_0()
