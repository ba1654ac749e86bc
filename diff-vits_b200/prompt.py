"""Native PromptEncoder (reference: model3.py:382-433): the speech-prompt encoder that produces the UNet's
encoder_hidden_states, run by libdiffvits_b200 (csrc/prompt_encoder.cu) on the engine's tcgen05 GEMM / attention / LayerNorm
kernels. The reference evaluates it inside every denoiser call (model3.py:906); it depends on the prompt only, so the drop-in
evaluates it once per job. Parameters stay in the reference's module (`Diffusion_Encoder.prompt_encoder`, same state_dict
keys); this class packs them and owns the plan for one (B, Tp)."""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import PromptCfg, check
from .engine import _ptr, _stream, pack_params, require_cuda


def supported(module) -> Optional[PromptCfg]:
    """The plan configuration if `module` is a PromptEncoder of the shape Diff-VITS' diffusion encoder uses (model3.py:898:
    PromptEncoder(100, hidden, hidden, 4, 0.2): 128 hidden, 8-head EncSALayer with the 9-tap conv feed-forward, no `g`
    conditioning), else None (the caller then runs the module itself)."""
    try:
        sd = module.state_dict()
        hidden = int(module.hidden_size)
        n_layers = int(module.num_layers)
        w = sd["pre.conv.weight"]
        if hidden != 128 or w.shape[0] != 1 or w.shape[2] != 128 or w.shape[1] > 128 or hasattr(module, "g_proj"):
            return None
        if not getattr(module, "last_ln", False) or sd["out_proj.conv.weight"].shape != (1, 128, 128):
            return None
        for l in range(n_layers):
            p = f"layers.{l}.op."
            if sd[p + "self_attn.in_proj_weight"].shape != (384, 128) or (p + "self_attn.in_proj_bias") in sd:
                return None
            if (p + "self_attn.out_proj.bias") in sd or sd[p + "ffn.ffn_1.8.weight"].shape != (512, 128):
                return None
            if (p + "ffn.ffn_1.9.weight") in sd or module.layers[l].op.self_attn.num_heads != 8:
                return None
        return PromptCfg(int(w.shape[1]), hidden, n_layers, 8, 9, 4)
    except (KeyError, AttributeError, IndexError):
        return None


class NativePromptEncoder:
    """One plan of the native PromptEncoder at fixed (B, Tp) on the module's CUDA device."""

    def __init__(self, module: torch.nn.Module, cfg: PromptCfg, B: int, Tp: int, device,
                 params: Optional[torch.Tensor] = None):
        self.device = require_cuda(device)
        self.lib = _lib.load()
        self.B, self.Tp, self.cfg = int(B), int(Tp), cfg
        self._plan = C.c_void_p()
        check(self.lib.dv_prompt_plan_create(C.byref(cfg), self.B, self.Tp, C.byref(self._plan)))
        self.manifest = json.loads(self.lib.dv_prompt_plan_manifest_json(self._plan).decode())
        self.param_bytes = self.lib.dv_prompt_plan_param_bytes(self._plan)
        if params is None:
            params = pack_params(self.manifest, module.state_dict(), self.param_bytes).to(self.device)
        self.params = params
        self.workspace = torch.empty(self.lib.dv_prompt_plan_workspace_bytes(self._plan), dtype=torch.uint8, device=self.device)
        check(self.lib.dv_prompt_plan_bind(self._plan, _ptr(self.workspace), _ptr(self.params), _stream(self.device)))

    def __call__(self, prompt: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        """prompt [B, in_channels, Tp], lengths [B] -> encoder_hidden_states [B, Tp, 128] fp32, padding rows zero
        (= (prompt_encoder(prompt, lengths) * mask).transpose(1, 2), model3.py:905-912)."""
        assert prompt.shape == (self.B, self.cfg.in_channels, self.Tp), prompt.shape
        prompt = prompt.to(self.device, torch.float32).contiguous()
        lengths = lengths.to(self.device, torch.int32).clamp(max=self.Tp).contiguous()
        out = torch.empty(self.B, self.Tp, self.cfg.hidden, dtype=torch.float32, device=self.device)
        check(self.lib.dv_prompt_encode(self._plan, _ptr(prompt), _ptr(lengths), _ptr(out), _stream(self.device)))
        return out

    def close(self) -> None:
        if getattr(self, "_plan", None) is not None and self._plan.value:
            torch.cuda.synchronize(self.device)
            self.lib.dv_prompt_plan_destroy(self._plan)
            self._plan = C.c_void_p()
        self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CACHE: Dict[Tuple, NativePromptEncoder] = {}


def encode_prompt(module: torch.nn.Module, prompt: torch.Tensor, lengths: torch.Tensor) -> Optional[torch.Tensor]:
    """encoder_hidden_states for a reference PromptEncoder module through the native plan (cached per module version and
    shape, at most 8 plans), or None when the module's configuration is outside what is built."""
    cfg = supported(module)
    if cfg is None or prompt.device.type != "cuda":
        return None
    ver = tuple((p.data_ptr(), p._version) for p in module.parameters())
    key = (id(module), ver, prompt.shape[0], prompt.shape[2], str(prompt.device))
    enc = _CACHE.get(key)
    if enc is None:
        stale = [k for k in _CACHE if k[0] == id(module) and k[1] != ver]
        for k in stale:
            _CACHE.pop(k).close()
        while len(_CACHE) >= 8:
            _CACHE.pop(next(iter(_CACHE))).close()
        same = next((e for k, e in _CACHE.items() if k[0] == id(module) and k[4] == str(prompt.device)), None)
        enc = _CACHE[key] = NativePromptEncoder(module, cfg, prompt.shape[0], prompt.shape[2], prompt.device,
                                                params=None if same is None else same.params)
    return enc(prompt, lengths)
