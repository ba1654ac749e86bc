"""Host side of the B200-native denoiser: weight packing and the ctypes calls into libdiffvits_b200.

`DenoiserEngine` owns one plan (fixed B, T, Tp on one device), its workspace and a handle on the packed parameter
blob. PyTorch is used for device memory (`torch.empty`), streams and the two hoisted, step-invariant prompt ops
(`TextTimeEmbedding`); all per-step compute is enqueued by the C library.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DVError, UnetCfg, check

SAMPLER_DPMPP_2M = 0
SAMPLER_UNIPC_BH2 = 1
SAMPLE_NO_GRAPH = 1
SAMPLE_ROUNDTRIP = 2


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(
            f"diffvits_b200 runs on CUDA sm_100a devices only (got {device}); there is no CPU fallback path")
    if not torch.cuda.is_available():
        raise RuntimeError("diffvits_b200: no CUDA device is available and there is no CPU fallback path")
    return device


# ------------------------------------------------------------------------------------------------ weight packing
def _interleave64(t: torch.Tensor) -> torch.Tensor:
    """[2H, ...] (value rows then gate rows) -> 64-row value / 64-row gate blocks alternating (GEGLU tiles)."""
    half = t.shape[0] // 2
    u, g = t[:half], t[half:]
    ub = u.reshape(half // 64, 64, *t.shape[1:])
    gb = g.reshape(half // 64, 64, *t.shape[1:])
    return torch.stack([ub, gb], dim=1).reshape(t.shape)


def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
    if t.shape[0] == rows:
        return t
    out = t.new_zeros((rows,) + tuple(t.shape[1:]))
    out[: t.shape[0]] = t
    return out


def pack_slot(slot: dict, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Build one slot of the parameter blob from reference-layout tensors. Returns a flat CPU tensor of the slot's
    dtype. Recipes are defined by the plan builder (csrc/engine.cu)."""
    rec, shape = slot["recipe"], slot["shape"]
    w = [sd[k].detach().to("cpu", torch.float32) for k in slot["keys"]]
    if rec == "time_freqs":
        from .dropin.unet1d.embeddings import time_frequencies  # same fp32 op sequence as the reference
        out = time_frequencies(shape[0], 0, 10000)
    elif rec == "vec":
        out = _pad_rows(torch.cat([v.reshape(-1) for v in w]), shape[0])
    elif rec == "vec_sum":
        out = _pad_rows(sum(v.reshape(-1) for v in w), shape[0])
    elif rec == "vec_dup2":
        out = _pad_rows(torch.cat([w[0].reshape(-1)] * 2), shape[0])
    elif rec == "geglu_vec":
        out = _pad_rows(_interleave64(w[0].reshape(-1)), shape[0])
    elif rec == "linear":
        mats = [v.reshape(v.shape[0], -1) for v in w]          # Linear [N, K] or 1x1 conv [N, K, 1]
        out = _pad_rows(torch.cat(mats, dim=0), shape[0])
    elif rec == "geglu":
        out = _pad_rows(_interleave64(w[0]), shape[0])
    elif rec == "conv3":
        n_pad, k_tot, cin = shape
        cpad = (cin + 63) // 64 * 64
        cw = w[0]                                              # [Cout, Cin, 3]
        assert cw.shape[1] == cin and cw.shape[2] == 3, (slot["name"], cw.shape)
        taps = cw.new_zeros(cw.shape[0], 3, cpad)
        taps[:, :, :cin] = cw.permute(0, 2, 1)                 # tap-major K: k = tap * cpad + c
        parts = [taps.reshape(cw.shape[0], 3 * cpad)]
        if len(w) > 1:                                         # 1x1 shortcut fused as extra K panels
            parts.append(w[1].reshape(w[1].shape[0], -1))
        out = _pad_rows(torch.cat(parts, dim=1), n_pad)
        assert out.shape[1] == k_tot, (slot["name"], out.shape, k_tot)
    elif rec in ("conv3_up2", "conv3_up2_odd"):
        # nearest resize folded into the conv (resnet.py:138-173), over the low-resolution rows p:
        #   to 2L rows,     K segments (x[p-1], x[p], x[p+1]):        even row 2p [W0 | W1+W2 | 0]   odd row 2p+1 [0 | W0+W1 | W2]
        #   to 2L - 1 rows, K segments (x[p-1], x[p], x'[p], x[p+1]): even [W0 | W1 | W2 | 0]         odd [0 | W0+W1 | 0 | W2]
        #   (x' = x without its last row: the W2 tap of the last even row falls on the conv padding)
        cw = w[0]
        w0, w1, w2 = cw[:, :, 0], cw[:, :, 1], cw[:, :, 2]
        z = torch.zeros_like(w0)
        if rec == "conv3_up2":
            even, odd = torch.cat([w0, w1 + w2, z], dim=1), torch.cat([z, w0 + w1, w2], dim=1)
        else:
            even, odd = torch.cat([w0, w1, w2, z], dim=1), torch.cat([z, w0 + w1, z, w2], dim=1)
        out = _pad_rows(torch.cat([even, odd], dim=0), shape[0])
    elif rec in ("convtbc1", "convtbc1_pad"):
        # ConvTBC k = 1 (model.py:60-70): weight [1, c_in, c_out] -> Linear layout [c_out, c_in], K zero-padded to the slot
        m = w[0][0].t().contiguous()
        out = m.new_zeros(shape[0], shape[1])
        out[: m.shape[0], : m.shape[1]] = m
    elif rec == "ffn_taps":
        # 9-tap conv feed-forward (operations.py:659-684): K = [tap 0 | ... | tap 8], the reference's kernel_size^-1/2 folded in
        out = _pad_rows(torch.cat(w, dim=1), shape[0]) * (len(w) ** -0.5)
    elif rec == "vec_ffn_scale":
        out = _pad_rows(w[0].reshape(-1), shape[0]) * (9 ** -0.5)
    elif rec == "mat_t":
        out = w[0].t().contiguous()
    elif rec == "rows":
        out = torch.cat([v.reshape(v.shape[0], -1) for v in w], dim=0)
    elif rec == "rows_t":
        out = torch.cat([v.reshape(v.shape[0], -1) for v in w], dim=0).t().contiguous()
    else:
        raise ValueError(f"unknown packing recipe {rec!r}")
    out = out.contiguous()
    n_expected = 1
    for d in (shape[:2] if rec == "conv3" else shape):
        n_expected *= d
    if out.numel() != n_expected:
        raise ValueError(f"slot {slot['name']}: packed {tuple(out.shape)} does not match manifest shape {shape}")
    if slot["dtype"] == "bf16x2":
        # FP32-class precision mode: the matrix as two bf16 planes, hi = bf16(w), lo = bf16(w - hi) (w = hi + lo to 16 bits)
        hi = out.to(torch.bfloat16)
        lo = (out - hi.to(torch.float32)).to(torch.bfloat16)
        return torch.cat([hi.reshape(-1), lo.reshape(-1)])
    if slot["dtype"] == "f16":
        return out.clamp(-65504.0, 65504.0).to(torch.float16).reshape(-1)
    return out.to(torch.bfloat16 if slot["dtype"] == "bf16" else torch.float32).reshape(-1)


def pack_params(manifest: List[dict], sd: Dict[str, torch.Tensor], nbytes: int) -> torch.Tensor:
    """Pack a reference-layout state_dict into the flat parameter blob (CPU uint8 tensor)."""
    blob = torch.zeros(nbytes, dtype=torch.uint8)
    for slot in manifest:
        flat = pack_slot(slot, sd)
        raw = flat.view(torch.uint8)
        assert raw.numel() == slot["nbytes"], (slot["name"], raw.numel(), slot["nbytes"])
        blob[slot["offset"]: slot["offset"] + slot["nbytes"]] = raw
    return blob


PRECISIONS = {"bf16": 0, "fp32x3": 1, "fp16s": 2, "fp16": 3}


def default_precision() -> str:
    """DV_PRECISION=bf16 (default, fast path) | fp16 (FP16 operands, the BF16 mode's speed, ~1e-3 per call, narrower range) |
    fp16s (FP16 operands + hi/lo residual stream, meets the 1e-3 gate) | fp32x3 (hi + lo BF16 operand split, three MMAs per
    product, FP32-class accuracy)."""
    import os
    p = os.environ.get("DV_PRECISION", "bf16").lower()
    if p not in PRECISIONS:
        raise ValueError(f"DV_PRECISION={p!r}: expected one of {sorted(PRECISIONS)}")
    return p


def unet_cfg_from_config(config: dict, x_channels: Optional[int] = None, precision: str = "bf16") -> UnetCfg:
    """Translate the reference UNet's config dict (unet_1d_condition.py:560-607) into the C struct; raises
    ValueError for configurations outside the accelerated path."""
    boc = tuple(config["block_out_channels"])
    n = len(boc)
    want_down = tuple(["CrossAttnDownBlock2D"] * (n - 1) + ["DownBlock2D"])
    want_up = tuple(["UpBlock2D"] + ["CrossAttnUpBlock2D"] * (n - 1))
    if tuple(config["down_block_types"]) != want_down or tuple(config["up_block_types"]) != want_up:
        raise ValueError("accelerated path supports the default CrossAttnDown*(n-1)+Down / Up+CrossAttnUp*(n-1) layout")
    lpb = config["layers_per_block"]
    if not isinstance(lpb, int):
        if len(set(lpb)) != 1:
            raise ValueError("layers_per_block must be uniform")
        lpb = lpb[0]
    heads = config["attention_head_dim"]
    if not isinstance(heads, int):
        if len(set(heads)) != 1:
            raise ValueError("attention_head_dim must be uniform")
        heads = heads[0]
    xdim = config["cross_attention_dim"]
    if not isinstance(xdim, int):
        if len(set(xdim)) != 1:
            raise ValueError("cross_attention_dim must be uniform")
        xdim = xdim[0]
    if config.get("addition_embed_type") not in ("text", None):
        raise ValueError("addition_embed_type must be 'text' or None")
    if config.get("resnet_time_scale_shift") != "scale_shift":
        raise ValueError("resnet_time_scale_shift must be 'scale_shift'")
    cfg = UnetCfg()
    cfg.in_channels = config["in_channels"]
    cfg.out_channels = config["out_channels"]
    cfg.n_levels = n
    for i, c in enumerate(boc):
        cfg.block_out_channels[i] = c
    cfg.layers_per_block = lpb
    cfg.norm_num_groups = config["norm_num_groups"]
    cfg.cross_attention_dim = xdim
    cfg.num_heads = heads
    cfg.norm_eps = config["norm_eps"]
    cfg.x_channels = x_channels if x_channels is not None else min(config["out_channels"], config["in_channels"])
    cfg.precision = PRECISIONS[precision]
    return cfg


# ------------------------------------------------------------------------------------------------ engine
class DenoiserEngine:
    """One plan of the native UNet forward / sampling loop at fixed (B, T, Tp) on one CUDA device."""

    def __init__(self, unet: torch.nn.Module, B: int, T: int, Tp: int, device, x_channels: Optional[int] = None,
                 params: Optional[torch.Tensor] = None, precision: str = "bf16"):
        self.device = require_cuda(device)
        self.lib = _lib.load()
        self.unet = unet
        self.B, self.T, self.Tp = int(B), int(T), int(Tp)
        self.precision = precision
        self.cfg = unet_cfg_from_config(unet.config, x_channels, precision)
        self.x_channels = self.cfg.x_channels
        self.in_channels, self.out_channels = self.cfg.in_channels, self.cfg.out_channels
        self._plan = C.c_void_p()
        check(self.lib.dv_plan_create(C.byref(self.cfg), self.B, self.T, self.Tp, C.byref(self._plan)))
        self.manifest = json.loads(self.lib.dv_plan_manifest_json(self._plan).decode())
        self.param_bytes = self.lib.dv_plan_param_bytes(self._plan)
        self.ws_bytes = self.lib.dv_plan_workspace_bytes(self._plan)
        self.launches_per_forward = self.lib.dv_plan_launches_per_forward(self._plan)
        with torch.cuda.device(self.device):
            if params is None:
                params = pack_params(self.manifest, unet.state_dict(), self.param_bytes).to(self.device)
            assert params.numel() == self.param_bytes and params.device == self.device
            self.params = params
            self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
            check(self.lib.dv_plan_bind(self._plan, _ptr(self.workspace), _ptr(self.params), _stream(self.device)))
        self._keep: list = []

    def close(self) -> None:
        """Destroy the plan (its graph, streams, events) and drop the workspace. Idempotent."""
        if getattr(self, "_plan", None) is not None and self._plan.value:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            self.lib.dv_plan_destroy(self._plan)
            self._plan = C.c_void_p()
        self.workspace = None
        self._keep = []
        self._keep_cond = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- step-invariant inputs
    def prepare_prompt(self, ehs: torch.Tensor, kv_len: Optional[torch.Tensor] = None,
                       aug: Optional[torch.Tensor] = None) -> None:
        """ehs: encoder_hidden_states [B, Tp, D]; kv_len: [B] valid prompt rows (None = no mask); aug: optional
        precomputed text-time embedding [B, 512] (default: computed here with the module's TextTimeEmbedding)."""
        assert ehs.shape == (self.B, self.Tp, self.cfg.cross_attention_dim), ehs.shape
        # The C library runs every plan call on the plan's device (api.cu DevGuard); the torch ops here (casts, the hoisted
        # TextTimeEmbedding) follow their tensors' device, and streams are looked up for self.device explicitly.
        ehs = ehs.to(self.device, torch.float32).contiguous()
        add = getattr(self.unet, "add_embedding", None)
        if aug is not None:
            aug = aug.to(self.device, torch.float32).contiguous()
            assert aug.shape == (self.B, 4 * self.cfg.block_out_channels[0])
        elif add is not None:
            with torch.no_grad(), torch.cuda.device(self.device):
                aug = add(ehs).to(torch.float32).contiguous()      # hoisted: prompt-only (unet_1d_condition.py:869)
        if kv_len is not None:
            kv_len = kv_len.to(self.device, torch.int32).contiguous()
        check(self.lib.dv_prompt_prepare(self._plan, _ptr(ehs), _ptr(aug), _ptr(kv_len), _stream(self.device)))
        self._keep = [ehs, aug, kv_len]

    def set_cond(self, cond: torch.Tensor) -> None:
        """cond: NCT [B, in_channels - x_channels, T] (model3.py:908)."""
        assert cond.shape == (self.B, self.in_channels - self.x_channels, self.T), cond.shape
        cond = cond.to(self.device, torch.float32).contiguous()
        check(self.lib.dv_set_cond(self._plan, _ptr(cond), _stream(self.device)))
        self._keep_cond = cond

    # ---- one denoiser call
    def forward(self, sample: torch.Tensor, t: torch.Tensor, x_only: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        cin = self.x_channels if x_only else self.in_channels
        assert sample.shape == (self.B, cin, self.T), (sample.shape, (self.B, cin, self.T))
        sample = sample.to(self.device, torch.float32).contiguous()
        t = t.to(self.device, torch.float32).reshape(-1).contiguous()
        assert t.numel() in (1, self.B)
        if out is None:
            out = torch.empty(self.B, self.out_channels, self.T, dtype=torch.float32, device=self.device)
        check(self.lib.dv_unet_forward(self._plan, _ptr(sample), 1 if x_only else 0, _ptr(t),
                                       1 if t.numel() == self.B else 0, _ptr(out), _stream(self.device)))
        return out

    # ---- whole sampling loop
    def sample(self, sampler: int, t_in: torch.Tensor, coef: torch.Tensor, x: torch.Tensor, graph: bool = True,
               roundtrip: bool = False) -> torch.Tensor:
        """x: NCT [B, x_channels, T] fp32 on the device, updated in place and returned."""
        assert x.shape == (self.B, self.x_channels, self.T) and x.dtype == torch.float32 and x.is_contiguous()
        assert x.device == self.device
        t_host = t_in.detach().to("cpu", torch.float32).contiguous()
        c_host = coef.detach().to("cpu", torch.float32).contiguous()
        nsteps = t_host.numel()
        assert c_host.shape == (nsteps, 8)
        flags = (0 if graph else SAMPLE_NO_GRAPH) | (SAMPLE_ROUNDTRIP if roundtrip else 0)
        fp = C.POINTER(C.c_float)
        check(self.lib.dv_sample(self._plan, sampler, nsteps, C.cast(t_host.data_ptr(), fp),
                                 C.cast(c_host.data_ptr(), fp), _ptr(x), flags, _stream(self.device)))
        return x

    def sample_diag(self) -> torch.Tensor:
        """Sampler diagnostics of the most recent `sample`: fp32 [rows, 4] view into the workspace (device), row j = j-th
        sampler-update launch: {non-finite count, sum of squares of the denoiser output it consumed, sum of squares of
        the state it wrote, kernel 0 DPM++ / 1 UniPC predictor / 2 UniPC corrector}. Reduced on the device inside the
        update kernels (dv_plan_sample_diag); reading it is the only synchronising step."""
        off, rows = C.c_size_t(), C.c_int()
        check(self.lib.dv_plan_sample_diag(self._plan, C.byref(off), C.byref(rows)))
        return self.workspace[off.value: off.value + rows.value * 16].view(torch.float32).view(rows.value, 4)

    def health_flag(self) -> torch.Tensor:
        """uint32 view into the workspace (device): non-zero once a bounded in-kernel wait gave up (dv_plan_health_flag)."""
        off = C.c_size_t()
        check(self.lib.dv_plan_health_flag(self._plan, C.byref(off)))
        return self.workspace[off.value: off.value + 4].view(torch.int32)

    def check_finite(self) -> None:
        """The reference asserts `torch.isnan(x).any() == False` on every denoiser call (model3.py:903, one device sync per
        step); here the update kernels count non-finite values and one read after the loop raises the same AssertionError."""
        _raise_nonfinite([self.sample_diag()], [self.health_flag()])

    # ---- debug: block outputs left in the workspace by the last denoiser call
    def named_tensors(self) -> Dict[str, torch.Tensor]:
        """{reference module path: bf16 [B, L, C] view into the workspace}."""
        out = {}
        buf = C.create_string_buffer(160)
        off, L, Cc = C.c_size_t(), C.c_int(), C.c_int()
        for i in range(self.lib.dv_plan_tensor_count(self._plan)):
            check(self.lib.dv_plan_tensor_info(self._plan, i, buf, 160, C.byref(off), C.byref(L), C.byref(Cc)))
            n = self.B * L.value * Cc.value * 2
            dt = torch.float16 if self.precision in ("fp16s", "fp16") else torch.bfloat16
            t = self.workspace[off.value: off.value + n].view(dt).view(self.B, L.value, Cc.value)
            if self.precision in ("fp32x3", "fp16s"):      # hi + lo planes -> the fp32 value they carry
                lo = self.workspace[off.value + n: off.value + 2 * n].view(dt).view(self.B, L.value, Cc.value)
                t = t.float() + lo.float()
            out[buf.value.decode()] = t
        return out

    # ---- measurement support
    def op_table(self) -> List[dict]:
        """[{tag, kind, flops, bytes}] for every launch of one denoiser call (kind: 0 gemm, 1 gn_stats, 2 gn_apply,
        3 layernorm, 4 attention)."""
        n = self.lib.dv_plan_op_count(self._plan)
        out = []
        buf = C.create_string_buffer(160)
        kind, fl, by = C.c_int(), C.c_double(), C.c_double()
        for i in range(n):
            check(self.lib.dv_plan_op_info(self._plan, i, buf, 160, C.byref(kind), C.byref(fl), C.byref(by)))
            out.append(dict(tag=buf.value.decode(), kind=kind.value, flops=fl.value, bytes=by.value))
        return out

    def profile_forward(self, sample: torch.Tensor, t: torch.Tensor, x_only: bool = True) -> List[float]:
        """Device time (ms) of every launch of one eager denoiser call, CUDA events on the launching stream;
        last entry = prologue (input layout change + time embedding). Synchronises."""
        n = self.lib.dv_plan_op_count(self._plan)
        ms = (C.c_float * (n + 1))()
        sample = sample.to(self.device, torch.float32).contiguous()
        t = t.to(self.device, torch.float32).reshape(-1).contiguous()
        out = torch.empty(self.B, self.out_channels, self.T, dtype=torch.float32, device=self.device)
        check(self.lib.dv_plan_profile_forward(self._plan, _ptr(sample), 1 if x_only else 0, _ptr(t),
                                               1 if t.numel() == self.B else 0, _ptr(out), ms, n + 1,
                                               _stream(self.device)))
        return list(ms)

    @property
    def last_sample_launches(self) -> int:
        return self.lib.dv_plan_last_sample_launches(self._plan)


def _raise_nonfinite(diags, flags=()) -> None:
    if flags and int(torch.stack([f.reshape(()) for f in flags]).sum()) != 0:
        raise RuntimeError("diffvits_b200: an in-kernel wait of the fused GroupNorm gave up (CTAs of one launch were not "
                           "co-scheduled); the results of this plan are invalid. Rebind the plan; DV_GN_OUT=0 disables the fusion")
    if not diags or diags[0].numel() == 0:
        return
    bad = torch.stack([d[:, 0] for d in diags]).sum(0).cpu()      # one small D2H (<= 2 KB) after the loop
    if bool((bad != 0).any()) or bool(torch.isnan(bad).any()):
        j = int(torch.nonzero((bad != 0) | torch.isnan(bad))[0])
        raise AssertionError(f"non-finite values in the denoising loop, first seen by sampler update {j} "
                             f"({int(bad[j]) if bad[j] == bad[j] else 'nan'} values); the reference asserts "
                             "torch.isnan(x).any() == False per call (model3.py:903)")


# ------------------------------------------------------------------------------------------------ lanes
def default_lanes(B: int) -> int:
    """How many independent batch slices ("lanes") the whole-loop graph runs on concurrent streams. Utterances never
    interact on this path, so a batch can be cut anywhere and lanes overlap one slice's launch gaps / kernel tails with
    another slice's kernels. Measured on C2 (B200): with the current kernels one lane (full-batch launches, most
    tiles per launch) is fastest: 220 ms vs 228 ms for two lanes; DV_LANES overrides."""
    import os
    env = os.environ.get("DV_LANES")
    if env:
        return max(1, min(int(env), B))
    return 1


class DenoiserGroup:
    """`lanes` DenoiserEngine plans over disjoint slices of one batch, sharing one packed parameter blob. Same
    interface as DenoiserEngine; `sample` captures all lanes into one CUDA graph (dv_sample_group)."""

    def __init__(self, unet: torch.nn.Module, B: int, T: int, Tp: int, device, x_channels: Optional[int] = None,
                 params: Optional[torch.Tensor] = None, lanes: Optional[int] = None, precision: str = "bf16"):
        self.precision = precision
        lanes = default_lanes(B) if lanes is None else max(1, min(int(lanes), B))
        base, extra = divmod(B, lanes)
        self.sizes = [base + (1 if i < extra else 0) for i in range(lanes)]
        self.bounds = [0]
        for n in self.sizes:
            self.bounds.append(self.bounds[-1] + n)
        self.engines: List[DenoiserEngine] = []
        for n in self.sizes:
            eng = DenoiserEngine(unet, n, T, Tp, device, x_channels=x_channels, params=params, precision=precision)
            params = eng.params
            self.engines.append(eng)
        e0 = self.engines[0]
        self.lib, self.device, self.params, self.cfg = e0.lib, e0.device, e0.params, e0.cfg
        self.B, self.T, self.Tp = int(B), int(T), int(Tp)
        self.x_channels, self.in_channels, self.out_channels = e0.x_channels, e0.in_channels, e0.out_channels
        self.manifest, self.param_bytes = e0.manifest, e0.param_bytes
        self.ws_bytes = sum(e.ws_bytes for e in self.engines)
        self.launches_per_forward = sum(e.launches_per_forward for e in self.engines)

    @property
    def lanes(self) -> int:
        return len(self.engines)

    def close(self) -> None:
        for e in self.engines:
            e.close()

    def _slices(self, t: Optional[torch.Tensor]):
        if t is None:
            return [None] * self.lanes
        return [t[self.bounds[i]: self.bounds[i + 1]] for i in range(self.lanes)]

    def prepare_prompt(self, ehs, kv_len=None, aug=None) -> None:
        assert ehs.shape[0] == self.B
        for e, a, b, c in zip(self.engines, self._slices(ehs), self._slices(kv_len), self._slices(aug)):
            e.prepare_prompt(a, b, c)

    def set_cond(self, cond) -> None:
        assert cond.shape[0] == self.B
        for e, c in zip(self.engines, self._slices(cond)):
            e.set_cond(c)

    def forward(self, sample, t, x_only: bool = False, out=None):
        assert sample.shape[0] == self.B
        t = t.reshape(-1)
        if out is None:
            out = torch.empty(self.B, self.out_channels, self.T, dtype=torch.float32, device=self.device)
        ts = self._slices(t) if t.numel() == self.B else [t] * self.lanes
        for e, s, tt, o in zip(self.engines, self._slices(sample), ts, self._slices(out)):
            e.forward(s, tt, x_only=x_only, out=o)
        return out

    def sample(self, sampler: int, t_in, coef, x, graph: bool = True, roundtrip: bool = False):
        assert x.shape == (self.B, self.x_channels, self.T) and x.dtype == torch.float32 and x.is_contiguous()
        assert x.device == self.device
        t_host = t_in.detach().to("cpu", torch.float32).contiguous()
        c_host = coef.detach().to("cpu", torch.float32).contiguous()
        nsteps = t_host.numel()
        assert c_host.shape == (nsteps, 8)
        flags = (0 if graph else SAMPLE_NO_GRAPH) | (SAMPLE_ROUNDTRIP if roundtrip else 0)
        fp = C.POINTER(C.c_float)
        n = self.lanes
        plans = (C.c_void_p * n)(*[e._plan for e in self.engines])
        xs = (C.c_void_p * n)(*[C.c_void_p(s.data_ptr()) for s in self._slices(x)])
        check(self.lib.dv_sample_group(plans, n, sampler, nsteps, C.cast(t_host.data_ptr(), fp),
                                       C.cast(c_host.data_ptr(), fp), xs, flags, _stream(self.device)))
        return x

    def sample_diag(self) -> torch.Tensor:
        """Diagnostics rows of lane 0 with the counts / sums of all lanes added (see DenoiserEngine.sample_diag)."""
        d = [e.sample_diag() for e in self.engines]
        out = d[0].clone()
        for o in d[1:]:
            out[:, :3] += o[:, :3]
        return out

    def check_finite(self) -> None:
        _raise_nonfinite([e.sample_diag() for e in self.engines], [e.health_flag() for e in self.engines])

    def named_tensors(self) -> Dict[str, torch.Tensor]:
        outs = [e.named_tensors() for e in self.engines]
        return {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}

    # measurement support: per-launch figures are those of lane 0 (the kernels as they run in the product)
    def op_table(self):
        return self.engines[0].op_table()

    def profile_forward(self, sample, t, x_only: bool = True):
        return self.engines[0].profile_forward(sample[: self.sizes[0]], t if t.numel() != self.B else t[: self.sizes[0]],
                                               x_only=x_only)

    @property
    def last_sample_launches(self) -> int:
        return self.engines[0].last_sample_launches
