"""Block containers of the 1-D UNet (reference: unet1d/unet_1d_blocks.py; the five classes the Diff-VITS config
instantiates: CrossAttnDownBlock2D :861, DownBlock2D :1019, UNetMidBlock2DCrossAttn :516, CrossAttnUpBlock2D :1986,
UpBlock2D :2134). Sub-modules are created in the reference's order (resnet, attention interleaved)."""
from torch import nn

from ._holder import ParamHolder
from .resnet import Downsample2D, ResnetBlock2D, Upsample2D
from .transformer_1d import Transformer2DModel


def _resnet(cin, cout, temb, eps, groups, scale_shift):
    return ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=temb, eps=eps, groups=groups,
                         time_embedding_norm=scale_shift)


def _attn(heads, channels, xdim, groups):
    return Transformer2DModel(heads, channels // heads, in_channels=channels, num_layers=1,
                              cross_attention_dim=xdim, norm_num_groups=groups)


class CrossAttnDownBlock2D(ParamHolder):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, resnet_eps, resnet_groups,
                 num_attention_heads, cross_attention_dim, add_downsample, resnet_time_scale_shift):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            resnets.append(_resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels, resnet_eps,
                                   resnet_groups, resnet_time_scale_shift))
            attentions.append(_attn(num_attention_heads, out_channels, cross_attention_dim, resnet_groups))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=1, name="op")]) if add_downsample else None)


class DownBlock2D(ParamHolder):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, resnet_eps, resnet_groups,
                 add_downsample, resnet_time_scale_shift):
        super().__init__()
        self.resnets = nn.ModuleList([
            _resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels, resnet_eps, resnet_groups,
                    resnet_time_scale_shift) for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=1, name="op")]) if add_downsample else None)


class UNetMidBlock2DCrossAttn(ParamHolder):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, resnet_eps, resnet_groups, num_attention_heads,
                 cross_attention_dim, resnet_time_scale_shift, num_layers=1):
        super().__init__()
        resnets = [_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups, resnet_time_scale_shift)]
        attentions = []
        for _ in range(num_layers):
            attentions.append(_attn(num_attention_heads, in_channels, cross_attention_dim, resnet_groups))
            resnets.append(_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups,
                                   resnet_time_scale_shift))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)


class CrossAttnUpBlock2D(ParamHolder):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, resnet_eps,
                 resnet_groups, num_attention_heads, cross_attention_dim, add_upsample, resnet_time_scale_shift):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            skip = in_channels if (i == num_layers - 1) else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(rin + skip, out_channels, temb_channels, resnet_eps, resnet_groups,
                                   resnet_time_scale_shift))
            attentions.append(_attn(num_attention_heads, out_channels, cross_attention_dim, resnet_groups))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)


class UpBlock2D(ParamHolder):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, resnet_eps,
                 resnet_groups, add_upsample, resnet_time_scale_shift):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_channels if (i == num_layers - 1) else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(rin + skip, out_channels, temb_channels, resnet_eps, resnet_groups,
                                   resnet_time_scale_shift))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
