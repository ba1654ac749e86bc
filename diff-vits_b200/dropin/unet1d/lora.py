"""LoRA-compatible layer types (reference: unet1d/lora.py:86-123). lora_layer is always None on this path, so they
are plain Conv1d / Linear parameter holders; the type names are kept because checkpoints and user code see them."""
from torch import nn


class LoRACompatibleConv(nn.Conv1d):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def set_lora_layer(self, lora_layer):
        if lora_layer is not None:
            raise NotImplementedError("LoRA layers are outside the accelerated Diff-VITS denoiser path")
        self.lora_layer = None


class LoRACompatibleLinear(nn.Linear):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def set_lora_layer(self, lora_layer):
        if lora_layer is not None:
            raise NotImplementedError("LoRA layers are outside the accelerated Diff-VITS denoiser path")
        self.lora_layer = None
