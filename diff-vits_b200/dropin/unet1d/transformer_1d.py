"""Transformer2DModel parameter holder (reference: unet1d/transformer_1d.py:41-326, continuous-input branch):
GroupNorm(eps 1e-6) -> 1x1 conv -> [B, L, C] -> BasicTransformerBlock(s) -> 1x1 conv -> + residual."""
from typing import Optional

from torch import nn

from ._holder import ParamHolder
from .attention import BasicTransformerBlock
from .lora import LoRACompatibleConv
from .outputs import BaseOutput


class Transformer2DModelOutput(BaseOutput):
    pass


class Transformer2DModel(ParamHolder):
    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: Optional[int] = None,
                 out_channels: Optional[int] = None, num_layers: int = 1, dropout: float = 0.0,
                 norm_num_groups: int = 32, cross_attention_dim: Optional[int] = None, attention_bias: bool = False,
                 sample_size=None, num_vector_embeds=None, patch_size=None, activation_fn: str = "geglu",
                 num_embeds_ada_norm=None, use_linear_projection: bool = False, only_cross_attention: bool = False,
                 upcast_attention: bool = False, norm_type: str = "layer_norm", norm_elementwise_affine: bool = True):
        super().__init__()
        if (in_channels is None or patch_size is not None or num_vector_embeds is not None or use_linear_projection
                or num_layers != 1):
            raise ValueError("only the continuous-input, conv-projection, single-layer Transformer2DModel is supported")
        self.use_linear_projection = use_linear_projection
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = LoRACompatibleConv(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  num_embeds_ada_norm=num_embeds_ada_norm, attention_bias=attention_bias,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                                  norm_type=norm_type, norm_elementwise_affine=norm_elementwise_affine)
            for _ in range(num_layers)
        ])
        self.out_channels = in_channels if out_channels is None else out_channels
        self.proj_out = LoRACompatibleConv(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)
