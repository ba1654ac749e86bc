"""UNet1DConditionModel - drop-in for adelacvg/diff-vits `unet1d/unet_1d_condition.py` (class :61, ctor :151-607,
forward :743-1037) whose forward pass is executed by libdiffvits_b200 on a B200.

Kept from the reference: import path, class name, constructor keyword arguments and defaults, the `config` dict,
the parameter names/shapes (`state_dict()` is key-for-key identical, so `load_state_dict(ckpt['model'])` of a
reference checkpoint works), the `forward(...)` signature and the `.sample` result attribute.
Not kept: a PyTorch execution path. `forward` requires a CUDA sm_100a device and raises otherwise.
"""
from __future__ import annotations

import collections
import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from .embeddings import TextTimeEmbedding, TimestepEmbedding, Timesteps
from .outputs import BaseOutput
from .unet_1d_blocks import (CrossAttnDownBlock2D, CrossAttnUpBlock2D, DownBlock2D, UNetMidBlock2DCrossAttn,
                             UpBlock2D)


class UNet1DConditionOutput(BaseOutput):
    """`.sample`: [B, out_channels, T] (reference: unet_1d_condition.py:48-58)."""


_CTOR_DEFAULTS = dict(
    sample_size=None, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn",
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
    mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280,
    transformer_layers_per_block=1, encoder_hid_dim=None, encoder_hid_dim_type=None, attention_head_dim=8,
    num_attention_heads=None, dual_cross_attention=False, use_linear_projection=False, class_embed_type=None,
    addition_embed_type=None, addition_time_embed_dim=None, num_class_embeds=None, upcast_attention=False,
    resnet_time_scale_shift="default", resnet_skip_time_act=False, resnet_out_scale_factor=1.0,
    time_embedding_type="positional", time_embedding_dim=None, time_embedding_act_fn=None, timestep_post_act=None,
    time_cond_proj_dim=None, conv_in_kernel=3, conv_out_kernel=3, projection_class_embeddings_input_dim=None,
    class_embeddings_concat=False, mid_block_only_cross_attention=None, cross_attention_norm=None,
    addition_embed_type_num_heads=64,
)

# options that must keep their default value on the accelerated path (everything model3.py never sets)
_FIXED = ("sample_size", "center_input_sample", "flip_sin_to_cos", "freq_shift", "only_cross_attention",
          "downsample_padding", "mid_block_scale_factor", "act_fn", "transformer_layers_per_block", "encoder_hid_dim",
          "encoder_hid_dim_type", "num_attention_heads", "dual_cross_attention", "use_linear_projection",
          "class_embed_type", "addition_time_embed_dim", "num_class_embeds", "upcast_attention",
          "resnet_skip_time_act", "resnet_out_scale_factor", "time_embedding_type", "time_embedding_dim",
          "time_embedding_act_fn", "timestep_post_act", "time_cond_proj_dim", "conv_in_kernel", "conv_out_kernel",
          "projection_class_embeddings_input_dim", "class_embeddings_concat", "mid_block_only_cross_attention",
          "cross_attention_norm", "mid_block_type")


def _accelerated_config(cfg: dict) -> Optional[str]:
    """None if the native engine implements this configuration, else the reason (mirrors Plan::build, csrc/engine.cu)."""
    boc = tuple(cfg["block_out_channels"])
    heads = cfg["attention_head_dim"]
    if not isinstance(heads, int) or not isinstance(cfg["cross_attention_dim"], int):
        return "per-level attention_head_dim / cross_attention_dim"
    if boc[0] != 128:
        return "block_out_channels[0] must be 128 (time_embed_dim 512)"
    for c in boc:
        if c % 64 or c > 1024 or c % heads or (c // heads) not in (16, 32, 48, 64):
            return f"channel count {c} with {heads} heads (head dim must be 16/32/48/64, channels a multiple of 64)"
        if cfg["norm_num_groups"] is None or c % (16 * cfg["norm_num_groups"]):
            return "channels per GroupNorm group must be a multiple of 16"
    if cfg["cross_attention_dim"] % 64 or cfg["out_channels"] > 128:
        return "cross_attention_dim must be a multiple of 64 and out_channels <= 128"
    return None


def _reference_class():
    """The reference's own UNet1DConditionModel, if a second `unet1d` package (the reference checkout) is on sys.path.
    Used only for UNet instances OUTSIDE the accelerated configuration (model3.py builds small UNets for the duration
    predictor and posterior encoder, :292, :493 - SURVEY section 8f 'next'), so that `import model3` keeps working with
    the drop-in installed. The diffusion denoiser never takes this route."""
    import importlib.util
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if "_dv_reference_unet1d" in sys.modules:
        return sys.modules["_dv_reference_unet1d.unet_1d_condition"].UNet1DConditionModel
    for entry in sys.path:
        cand = os.path.join(entry or ".", "unet1d")
        f = os.path.join(cand, "unet_1d_condition.py")
        if os.path.isfile(f) and os.path.abspath(cand) != here:
            spec = importlib.util.spec_from_file_location("_dv_reference_unet1d", os.path.join(cand, "__init__.py"),
                                                          submodule_search_locations=[cand])
            pkg = importlib.util.module_from_spec(spec)
            sys.modules["_dv_reference_unet1d"] = pkg
            spec.loader.exec_module(pkg)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mod = importlib.import_module("_dv_reference_unet1d.unet_1d_condition")
            return mod.UNet1DConditionModel
    return None


class UNet1DConditionModel(nn.Module):
    def __new__(cls, *args, **kwargs):
        # copy.deepcopy / pickle re-create instances through cls.__new__(cls) without arguments: never delegate then
        if cls is UNet1DConditionModel and (args or kwargs) and not (set(kwargs) - set(_CTOR_DEFAULTS)):
            cfg = dict(_CTOR_DEFAULTS)
            cfg.update(dict(zip(list(_CTOR_DEFAULTS), args)))
            cfg.update({k: v for k, v in kwargs.items() if k in _CTOR_DEFAULTS})
            try:
                reason = _accelerated_config(cfg)
            except Exception:
                reason = None            # malformed arguments: let __init__ raise the proper error
            if reason is not None:
                ref = _reference_class()
                if ref is None:
                    raise ValueError(f"UNet1DConditionModel configuration outside the B200-accelerated path ({reason}) "
                                     "and no reference `unet1d` package on sys.path to delegate to")
                return ref(*args, **kwargs)          # not an instance of cls: our __init__ is skipped
        return super().__new__(cls)

    def __init__(self, *args, **kwargs):
        super().__init__()
        names = list(_CTOR_DEFAULTS)
        if len(args) > len(names):
            raise TypeError("too many positional arguments")
        for name, val in zip(names, args):
            if name in kwargs:
                raise TypeError(f"got multiple values for argument {name!r}")
            kwargs[name] = val
        unknown = set(kwargs) - set(_CTOR_DEFAULTS)
        if unknown:
            raise TypeError(f"UNet1DConditionModel got unexpected keyword arguments {sorted(unknown)}")
        cfg = dict(_CTOR_DEFAULTS)
        cfg.update(kwargs)
        for name in _FIXED:
            if cfg[name] != _CTOR_DEFAULTS[name]:
                raise ValueError(f"`{name}`={cfg[name]!r} is outside the accelerated Diff-VITS denoiser path "
                                 f"(only {_CTOR_DEFAULTS[name]!r} is implemented)")
        boc = tuple(cfg["block_out_channels"])
        n = len(boc)
        down_types, up_types = tuple(cfg["down_block_types"]), tuple(cfg["up_block_types"])
        # same argument checks (and messages) as the reference ctor, unet_1d_condition.py:222-255
        if len(down_types) != len(up_types):
            raise ValueError(f"Must provide the same number of `down_block_types` as `up_block_types`. "
                             f"`down_block_types`: {down_types}. `up_block_types`: {up_types}.")
        if len(boc) != len(down_types):
            raise ValueError(f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                             f"`block_out_channels`: {boc}. `down_block_types`: {down_types}.")
        if down_types != tuple(["CrossAttnDownBlock2D"] * (n - 1) + ["DownBlock2D"]) or \
                up_types != tuple(["UpBlock2D"] + ["CrossAttnUpBlock2D"] * (n - 1)):
            raise ValueError("only the default block layout (cross-attention everywhere but the deepest level) is "
                             "implemented on the accelerated path")
        if cfg["resnet_time_scale_shift"] != "scale_shift":
            raise ValueError("only resnet_time_scale_shift='scale_shift' is implemented (model3.py:895)")
        if cfg["addition_embed_type"] not in (None, "text"):
            raise ValueError(f"addition_embed_type: {cfg['addition_embed_type']} must be None or 'text'.")
        for key in ("layers_per_block", "attention_head_dim", "cross_attention_dim"):
            if not isinstance(cfg[key], int):
                raise ValueError(f"`{key}` must be an int on the accelerated path")
        if cfg["norm_num_groups"] is None:
            raise ValueError("norm_num_groups=None is not implemented")

        in_channels, out_channels = cfg["in_channels"], cfg["out_channels"]
        heads, xdim, groups, eps = cfg["attention_head_dim"], cfg["cross_attention_dim"], cfg["norm_num_groups"], cfg["norm_eps"]
        lpb, ss = cfg["layers_per_block"], cfg["resnet_time_scale_shift"]
        self.sample_size = cfg["sample_size"]

        # --- construction order == reference order, so seeded default init gives identical weights
        self.conv_in = nn.Conv1d(in_channels, boc[0], kernel_size=3, padding=1)
        time_embed_dim = boc[0] * 4
        self.time_proj = Timesteps(boc[0], cfg["flip_sin_to_cos"], cfg["freq_shift"])
        self.time_embedding = TimestepEmbedding(boc[0], time_embed_dim, act_fn=cfg["act_fn"])
        self.encoder_hid_proj = None
        self.class_embedding = None
        if cfg["addition_embed_type"] == "text":
            self.add_embedding = TextTimeEmbedding(xdim, time_embed_dim, num_heads=cfg["addition_embed_type_num_heads"])
        self.time_embed_act = None

        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        output_channel = boc[0]
        for i in range(n):
            input_channel, output_channel = output_channel, boc[i]
            final = i == n - 1
            if not final:
                blk = CrossAttnDownBlock2D(input_channel, output_channel, time_embed_dim, lpb, eps, groups, heads, xdim,
                                           add_downsample=True, resnet_time_scale_shift=ss)
            else:
                blk = DownBlock2D(input_channel, output_channel, time_embed_dim, lpb, eps, groups,
                                  add_downsample=False, resnet_time_scale_shift=ss)
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], time_embed_dim, eps, groups, heads, xdim, ss)
        self.num_upsamplers = 0
        rev = list(reversed(boc))
        output_channel = rev[0]
        for i in range(n):
            final = i == n - 1
            prev_output_channel, output_channel = output_channel, rev[i]
            input_channel = rev[min(i + 1, n - 1)]
            if not final:
                self.num_upsamplers += 1
            if i == 0:
                blk = UpBlock2D(input_channel, prev_output_channel, output_channel, time_embed_dim, lpb + 1, eps, groups,
                                add_upsample=not final, resnet_time_scale_shift=ss)
            else:
                blk = CrossAttnUpBlock2D(input_channel, output_channel, prev_output_channel, time_embed_dim, lpb + 1,
                                         eps, groups, heads, xdim, add_upsample=not final,
                                         resnet_time_scale_shift=ss)
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=groups, eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv1d(boc[0], out_channels, kernel_size=3, padding=1)

        self.config = cfg
        # shape -> plan (own workspace), least recently used first; bounded so that a long-running process that sees a new T
        # per utterance (ddpm / ddim sampling, validation) does not grow GPU memory without limit
        self._engines: "collections.OrderedDict[Tuple, Any]" = collections.OrderedDict()
        self._engine_cache_size = int(os.environ.get("DV_PLAN_CACHE", "4"))
        self._params_cache: Optional[Tuple[Tuple, torch.Tensor]] = None
        # arithmetic of the native path: "bf16" (fast path) or "fp32x3" (hi + lo BF16 operand split, FP32-class accuracy:
        # include/diffvits_b200.h DV_PRECISION_*). Not a constructor argument (the reference's ctor has none): set the
        # attribute, or DV_PRECISION in the environment, before the first forward / sample.
        from diffvits_b200.engine import default_precision
        self.native_precision = default_precision()

    # ------------------------------------------------------------------------------------------ engine plumbing
    @property
    def device(self) -> torch.device:
        return self.conv_in.weight.device

    @property
    def dtype(self) -> torch.dtype:
        return self.conv_in.weight.dtype

    def _weights_version(self) -> Tuple:
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def native_engine(self, B: int, T: int, Tp: int, x_channels: Optional[int] = None, private: bool = False,
                      lanes: Optional[int] = None):
        """The (cached) native plan for this shape on the module's device. Weights are packed once per module and
        shared by all of its plans; they are re-packed when any parameter changed (load_state_dict, .to(...)).
        `private=True` returns a fresh plan (own workspace, prompt and cond state) that is not shared with
        `forward()` or other callers."""
        from diffvits_b200.engine import DenoiserGroup, require_cuda
        device = require_cuda(self.device)
        if self.dtype != torch.float32:
            raise RuntimeError("keep UNet1DConditionModel parameters in fp32; the engine packs BF16 operands itself")
        ver = (str(device), self.native_precision) + self._weights_version()
        if self._params_cache is not None and self._params_cache[0] != ver:
            self._params_cache = None
            for e in self._engines.values():
                e.close()
            self._engines.clear()          # plans hold pointers into the old blob
        key = (B, T, Tp, x_channels, lanes)
        eng = None if private else self._engines.get(key)
        if eng is None:
            blob = None if self._params_cache is None else self._params_cache[1]
            eng = DenoiserGroup(self, B, T, Tp, device, x_channels=x_channels, params=blob, lanes=lanes,
                                precision=self.native_precision)
            if self._params_cache is None:
                self._params_cache = (ver, eng.params)
            if not private:
                self._engines[key] = eng
                while len(self._engines) > max(1, self._engine_cache_size):
                    _, old = self._engines.popitem(last=False)      # least recently used plan: destroy it, free its workspace
                    old.close()
        elif not private:
            self._engines.move_to_end(key)
        return eng

    def native_denoiser(self, cond: torch.Tensor, ehs: torch.Tensor, kv_len: Optional[torch.Tensor] = None,
                        aug: Optional[torch.Tensor] = None):
        """A `NativeDenoiser` for this batch, re-using (rebinding) the cached one of the same shape: plan, workspace and the
        captured loop graph survive from one `solver.sample()` call to the next (tts_infer.py calls it once per utterance)."""
        from diffvits_b200.native import NativeDenoiser
        B, Cc, T = cond.shape
        key = ("denoiser", B, T, ehs.shape[1], Cc)
        ver = (str(self.device), self.native_precision) + self._weights_version()
        if self._params_cache is not None and self._params_cache[0] != ver:
            self._params_cache = None
            for e in self._engines.values():
                e.close()
            self._engines.clear()
        nd = self._engines.get(key)
        if nd is None:
            nd = NativeDenoiser(self, cond, ehs, kv_len, aug)
            self._engines[key] = nd
            while len(self._engines) > max(1, self._engine_cache_size):
                _, old = self._engines.popitem(last=False)
                old.close()
        else:
            self._engines.move_to_end(key)
            nd.rebind(cond, ehs, kv_len, aug)
        return nd

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, class_labels: Optional[torch.Tensor] = None,
                timestep_cond: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None,
                encoder_attention_mask: Optional[torch.Tensor] = None, return_dict: bool = True):
        """One denoiser call (reference: unet_1d_condition.py:743-1037).

        sample [B, in_channels, T]; timestep: python number, 0-d or [B] tensor (fractional model time);
        encoder_hidden_states [B, Tp, cross_attention_dim]; encoder_attention_mask: bool/0-1 [B, Tp] key-padding
        mask (True = keep) as `Diffusion_Encoder.forward` builds it (model3.py:911-912)."""
        for name, val in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                          ("attention_mask", attention_mask), ("cross_attention_kwargs", cross_attention_kwargs),
                          ("added_cond_kwargs", added_cond_kwargs),
                          ("down_block_additional_residuals", down_block_additional_residuals),
                          ("mid_block_additional_residual", mid_block_additional_residual)):
            if val is not None and not (isinstance(val, dict) and not val):
                raise NotImplementedError(f"`{name}` is not used by the Diff-VITS denoiser and not implemented natively")
        # Inference only: the native path has no backward. Running it where autograd expects a graph (the reference's
        # train.py / model3 training forward) would silently cut the denoiser out of the loss, so refuse instead.
        if torch.is_grad_enabled() and (self.training or sample.requires_grad or encoder_hidden_states.requires_grad):
            raise RuntimeError(
                "diffvits_b200 UNet1DConditionModel.forward is inference-only (no autograd through the native kernels): "
                "put the module in eval() mode and call it under torch.no_grad() (model3.py:1120 does), or use the reference "
                "unet1d package for training")
        if sample.dim() != 3 or encoder_hidden_states.dim() != 3:
            raise ValueError("expected sample [B, C, T] and encoder_hidden_states [B, Tp, D]")
        B, _, T = sample.shape
        Tp = encoder_hidden_states.shape[1]
        eng = self.native_engine(B, T, Tp)
        dev = eng.device
        kv_len = None
        if encoder_attention_mask is not None:
            kv_len = mask_to_lengths(encoder_attention_mask, Tp)
        with torch.no_grad():
            eng.prepare_prompt(encoder_hidden_states.to(dev), kv_len)
        if not torch.is_tensor(timestep):
            t = torch.tensor([float(timestep)], dtype=torch.float32, device=dev)
        else:
            t = timestep.to(dev, torch.float32).reshape(-1)
        out = eng.forward(sample.to(dev), t, x_only=False).to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNet1DConditionOutput(sample=out)


def mask_to_lengths(mask: torch.Tensor, Tp: int) -> torch.Tensor:
    """[B, Tp] keep-mask (a prefix mask, as commons.sequence_mask produces) -> [B] int32 lengths."""
    if mask.dim() != 2 or mask.shape[1] != Tp:
        raise NotImplementedError("encoder_attention_mask must be a [B, Tp] key-padding mask on the accelerated path")
    m = mask.to(torch.bool)
    lens = m.sum(dim=1)
    prefix = torch.arange(Tp, device=m.device)[None, :] < lens[:, None]
    if not bool((prefix == m).all()):
        raise NotImplementedError("only prefix (sequence_mask-style) prompt masks are implemented natively")
    return lens.to(torch.int32)
