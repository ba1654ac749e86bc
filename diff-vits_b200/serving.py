"""Batched / bucketed front of the denoising loop.

The reference synthesises ONE utterance per call (`tts_infer.py:46-74, 152-163`: batch 1, a new frame count T from the
duration predictor every time). Utterances never interact on this path, so requests can share launches - but only in a way
that leaves every request's result what the reference would have produced for it alone:

  * frames: convolutions and self-attention see every frame of the batch tensor (the reference has no frame mask,
    `model3.py:910` is commented out), so requests are batched only with requests of EXACTLY the same T (no frame padding);
  * prompt: cross-attention takes a key-padding mask (`model3.py:911-912`), the additive -10000 bias underflows to an exact
    zero weight, so prompts of different length CAN share a batch: they are zero-padded to the bucket's prompt length and
    masked by `kv_len`. The text-time embedding (`unet_1d_condition.py:869-870`: AttentionPooling attends over every prompt
    row it is given, padded or not) is therefore computed per request on its UNPADDED prompt and passed in (`aug`).

Buckets are keyed by (T, padded prompt length); each bucket runs as batches of at most `max_batch` through a cached
`NativeDenoiser` (plan, workspace and the captured loop graph are re-used whenever a (B, T, Tp) shape comes back).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch


@dataclass
class Request:
    """One utterance: cond [Cc, T] (prior output, model3.py:1173), ehs [Tp, D] (masked, transposed PromptEncoder output,
    model3.py:906, 912), noise [x_channels, T] or None (drawn from `generator`)."""
    cond: torch.Tensor
    ehs: torch.Tensor
    noise: Optional[torch.Tensor] = None


def prompt_bucket(tp: int, granularity: int = 64) -> int:
    """Padded prompt length of a request: the next multiple of `granularity` (fewer distinct plan shapes)."""
    return max(granularity, (tp + granularity - 1) // granularity * granularity)


def plan_batches(shapes: Sequence[Tuple[int, int]], max_batch: int, granularity: int = 64) -> List[Tuple[Tuple[int, int], List[int]]]:
    """shapes[i] = (T_i, Tp_i). Returns [((T, Tp_padded), [request indices])]: requests grouped by exact T and padded prompt
    length, in submission order inside a group, split into batches of at most `max_batch`. Pure host logic (unit-tested)."""
    groups: Dict[Tuple[int, int], List[int]] = {}
    for i, (t, tp) in enumerate(shapes):
        groups.setdefault((int(t), prompt_bucket(int(tp), granularity)), []).append(i)
    out = []
    for key in sorted(groups):
        idx = groups[key]
        for k in range(0, len(idx), max_batch):
            out.append((key, idx[k:k + max_batch]))
    return out


class BucketedSampler:
    """Runs a list of independent requests through the native sampling loop in shape buckets.

    unet: drop-in UNet1DConditionModel on a CUDA device; betas: the model's discrete schedule (model3.py:978);
    sampler / steps / order as in `NaturalSpeech2.sample` (model3.py:1143-1158 dpmsolver++, 1177-1192 unipc)."""

    def __init__(self, unet, betas: torch.Tensor, sampler: str = "unipc", steps: int = 30, order: int = 2,
                 max_batch: int = 32, prompt_granularity: int = 64, generator: Optional[torch.Generator] = None):
        from .engine import require_cuda
        self.unet, self.betas = unet, betas
        self.sampler, self.steps, self.order = sampler, int(steps), int(order)
        self.max_batch, self.gran = int(max_batch), int(prompt_granularity)
        self.device = require_cuda(unet.device)
        self.generator = generator
        self.stats = {"batches": 0, "requests": 0, "shapes": set()}

    @torch.no_grad()
    def run(self, requests: Sequence[Request]) -> List[torch.Tensor]:
        """Latents [x_channels, T_i] in submission order."""
        from .native import sample_native
        dev = self.device
        xch = self.unet.config["in_channels"] - requests[0].cond.shape[0] if requests else 0
        shapes = [(r.cond.shape[1], r.ehs.shape[0]) for r in requests]
        out: List[Optional[torch.Tensor]] = [None] * len(requests)
        for (T, Tp), idx in plan_batches(shapes, self.max_batch, self.gran):
            B = len(idx)
            cond = torch.stack([requests[i].cond for i in idx]).to(dev, torch.float32)
            D = requests[idx[0]].ehs.shape[1]
            ehs = torch.zeros(B, Tp, D, device=dev)
            kv_len = torch.empty(B, dtype=torch.int32, device=dev)
            aug = []
            for k, i in enumerate(idx):
                e = requests[i].ehs.to(dev, torch.float32)
                ehs[k, : e.shape[0]] = e
                kv_len[k] = e.shape[0]
                # AttentionPooling over exactly the rows the reference would see for this request alone
                aug.append(self.unet.add_embedding(e[None]).to(torch.float32))
            aug = torch.cat(aug, dim=0)
            noise = []
            for i in idx:
                n = requests[i].noise
                if n is None:
                    n = torch.randn(xch, T, device=dev, generator=self.generator)
                noise.append(n.to(dev, torch.float32))
            x_T = torch.stack(noise)
            native = self.unet.native_denoiser(cond, ehs, kv_len, aug)
            lat = sample_native(self.unet, x_T, cond, ehs, kv_len, sampler=self.sampler, steps=self.steps, order=self.order,
                                betas=self.betas, native=native)
            for k, i in enumerate(idx):
                out[i] = lat[k].clone()
            self.stats["batches"] += 1
            self.stats["requests"] += B
            self.stats["shapes"].add((B, T, Tp))
        return out  # type: ignore[return-value]
