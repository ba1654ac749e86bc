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


class PromptEncoderParams(torch.nn.Module):
    """Parameter container with the reference PromptEncoder's state_dict layout (model3.py:382-433; keys `pre.*`,
    `layers.N.op.*`, `out_proj.*`, `layer_norm.*`) for callers that do not have the reference's `model3` module at hand
    (bench.py, tests): `load_state_dict` of a reference checkpoint's `diff_model.prompt_encoder.*` entries works. No
    `forward`: compute is `encode_prompt` / `NativePromptEncoder`. Random init: LayerNorm (1, 0), normal weights."""

    class _Holder(torch.nn.Module):
        pass

    def __init__(self, in_channels: int = 100, hidden: int = 128, n_layers: int = 4, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        P = torch.nn.Parameter

        def ln(c):
            h = self._Holder()
            h.weight, h.bias = P(torch.ones(c)), P(torch.zeros(c))
            return h

        def randn(*shape, std):
            return P(torch.randn(*shape, generator=g) * std)

        self.hidden_size, self.num_layers, self.last_ln = hidden, n_layers, True
        self.pre = self._Holder()
        self.pre.layer_norm = ln(in_channels)
        self.pre.conv = self._Holder()
        self.pre.conv.weight, self.pre.conv.bias = randn(1, in_channels, hidden, std=(3.2 / in_channels) ** 0.5), P(torch.zeros(hidden))
        self.layers = torch.nn.ModuleList()
        for _ in range(n_layers):
            lay, op = self._Holder(), self._Holder()
            op.layer_norm1, op.layer_norm2 = ln(hidden), ln(hidden)
            op.self_attn = self._Holder()
            op.self_attn.num_heads = 8
            op.self_attn.in_proj_weight = randn(3 * hidden, hidden, std=hidden ** -0.5)
            op.self_attn.out_proj = self._Holder()
            op.self_attn.out_proj.weight = randn(hidden, hidden, std=hidden ** -0.5)
            op.ffn = self._Holder()
            op.ffn.ffn_1 = torch.nn.ModuleList()
            for i in range(9):
                lin = self._Holder()
                lin.weight = randn(4 * hidden, hidden, std=hidden ** -0.5)
                if i == 0:
                    lin.bias = P(torch.zeros(4 * hidden))
                op.ffn.ffn_1.append(lin)
            op.ffn.ffn_2 = self._Holder()
            op.ffn.ffn_2.weight, op.ffn.ffn_2.bias = randn(hidden, 4 * hidden, std=(4 * hidden) ** -0.5), P(torch.zeros(hidden))
            lay.op = op
            self.layers.append(lay)
        self.out_proj = self._Holder()
        self.out_proj.layer_norm = ln(hidden)
        self.out_proj.conv = self._Holder()
        self.out_proj.conv.weight, self.out_proj.conv.bias = randn(1, hidden, hidden, std=(4.0 / hidden) ** 0.5), P(torch.zeros(hidden))
        self.layer_norm = ln(hidden)

    def forward(self, *_a, **_k):
        raise RuntimeError("PromptEncoderParams holds parameters only; use diffvits_b200.prompt.encode_prompt (CUDA, no CPU fallback)")


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
