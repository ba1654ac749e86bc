"""BasicTransformerBlock / FeedForward / GEGLU parameter holders (reference: unet1d/attention.py:26-301).
Pre-norm residual block: x += attn1(LN1 x); x += attn2(LN2 x, prompt); x += FF(LN3 x), FF = Linear(GEGLU(x))."""
from typing import Optional

from torch import nn

from ._holder import ParamHolder
from .attention_processor import Attention
from .lora import LoRACompatibleLinear


class GEGLU(ParamHolder):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = LoRACompatibleLinear(dim_in, dim_out * 2)


class FeedForward(ParamHolder):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu", final_dropout: bool = False):
        super().__init__()
        if activation_fn != "geglu" or final_dropout:
            raise ValueError("only the GEGLU feed-forward of the Diff-VITS UNet is supported")
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        act = GEGLU(dim, inner_dim)
        self.net = nn.ModuleList([act, nn.Dropout(dropout), LoRACompatibleLinear(inner_dim, dim_out)])


class BasicTransformerBlock(ParamHolder):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, dropout=0.0,
                 cross_attention_dim: Optional[int] = None, activation_fn: str = "geglu",
                 num_embeds_ada_norm: Optional[int] = None, attention_bias: bool = False,
                 only_cross_attention: bool = False, double_self_attention: bool = False,
                 upcast_attention: bool = False, norm_elementwise_affine: bool = True, norm_type: str = "layer_norm",
                 final_dropout: bool = False):
        super().__init__()
        if (num_embeds_ada_norm is not None or only_cross_attention or double_self_attention or attention_bias
                or norm_type != "layer_norm" or not norm_elementwise_affine or cross_attention_dim is None):
            raise ValueError("unsupported BasicTransformerBlock variant for the Diff-VITS denoiser")
        self.only_cross_attention = only_cross_attention
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=True)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias, cross_attention_dim=None, upcast_attention=upcast_attention)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=True)
        self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                               dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                               upcast_attention=upcast_attention)
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=True)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)
