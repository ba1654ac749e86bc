"""ResnetBlock2D / Downsample2D / Upsample2D parameter holders (1-D despite the names).
Reference: unet1d/resnet.py:461-641, 176-223, 104-173. Construction order of the sub-layers follows the reference
so that seeded default initialisation produces identical weights."""
from torch import nn

from ._holder import ParamHolder
from .lora import LoRACompatibleConv, LoRACompatibleLinear


class Upsample2D(ParamHolder):
    """nearest resize to the skip's length, then conv k3 p1 (reference: resnet.py:138-173)."""

    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        if use_conv_transpose or not use_conv:
            raise ValueError("only the use_conv=True upsampler of the Diff-VITS UNet is supported")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv, self.use_conv_transpose, self.name = use_conv, use_conv_transpose, name
        self.conv = LoRACompatibleConv(self.channels, self.out_channels, 3, padding=1)


class Downsample2D(ParamHolder):
    """conv k3 s2 p1 (reference: resnet.py:176-223)."""

    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        if not use_conv or padding != 1:
            raise ValueError("only the use_conv=True, padding=1 downsampler of the Diff-VITS UNet is supported")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        self.conv = LoRACompatibleConv(self.channels, self.out_channels, 3, stride=2, padding=padding)
        if name == "conv":
            self.Conv2d_0 = self.conv


class ResnetBlock2D(ParamHolder):
    """GN -> SiLU -> conv3 -> GN -> (1+scale)*h+shift -> SiLU -> conv3, plus identity / 1x1 shortcut."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512,
                 groups=32, groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None,
                 up=False, down=False, conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        if time_embedding_norm != "scale_shift":
            raise ValueError("the accelerated path implements resnet_time_scale_shift='scale_shift' (model3.py:895)")
        if up or down or kernel is not None or skip_time_act or output_scale_factor != 1.0:
            raise ValueError("unsupported ResnetBlock2D variant for the Diff-VITS denoiser")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.time_embedding_norm = time_embedding_norm
        self.output_scale_factor = output_scale_factor
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = LoRACompatibleConv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = LoRACompatibleLinear(temb_channels, 2 * out_channels)
        self.norm2 = nn.GroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        conv_2d_out_channels = conv_2d_out_channels or out_channels
        self.conv2 = LoRACompatibleConv(out_channels, conv_2d_out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.upsample = self.downsample = None
        self.use_in_shortcut = in_channels != conv_2d_out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = None
        if self.use_in_shortcut:
            self.conv_shortcut = LoRACompatibleConv(in_channels, conv_2d_out_channels, kernel_size=1, stride=1,
                                                    padding=0, bias=conv_shortcut_bias)
