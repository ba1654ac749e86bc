"""Batched / bucketed serving front (diff-vits_b200/serving.py) on the B200: every request's latent equals what
the same request gives when it runs ALONE (the reference's batch-1 use, tts_infer.py:152-163)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bucketed_requests_equal_single_request_runs(seeded_unet):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import diffvits_b200
    import gpu_util as U
    from diffvits_b200.serving import BucketedSampler, Request
    from oracle import inputs
    unet = copy.deepcopy(seeded_unet).to("cuda")
    betas = inputs.betas_linear()
    g = torch.Generator().manual_seed(5)
    shapes = [(136, 40), (200, 282), (136, 100), (136, 64), (200, 150), (136, 40)]
    reqs = []
    for T, Tp in shapes:
        reqs.append(Request(cond=torch.randn(128, T, generator=g), ehs=torch.randn(Tp, 128, generator=g),
                            noise=torch.randn(100, T, generator=g)))
    front = BucketedSampler(unet, betas, sampler="unipc", steps=6, max_batch=2)
    outs = front.run(reqs)
    assert front.stats["requests"] == len(reqs) and front.stats["batches"] == 5      # (136,64) x 2 batches, (136,128), (200,192), (200,320)
    for i, (r, lat) in enumerate(zip(reqs, outs)):
        assert lat.shape == (100, r.cond.shape[1]) and torch.isfinite(lat).all()
        # the same request alone, prompt padded the same way: bit for bit (no cross-sample reduction anywhere)
        alone = BucketedSampler(unet, betas, sampler="unipc", steps=6, max_batch=1).run([r])[0]
        assert torch.equal(lat, alone), i
        # the same request alone with its UNPADDED prompt and no mask, as the reference runs it: same result up to the order
        # of the fp32 accumulation over key tiles (masked keys weigh exactly zero), i.e. the BF16 loop floor
        cond, ehs = r.cond[None].cuda(), r.ehs[None].cuda()
        ref = diffvits_b200.sample_native(unet, r.noise[None].cuda(), cond, ehs, sampler="unipc", steps=6, betas=betas)
        rel = U.rel_l2(lat[None], ref)
        print(f"request {i} (T={r.cond.shape[1]}, Tp={r.ehs.shape[0]}): bucketed vs unpadded single run rel {rel:.3e}")
        assert rel < 4e-2
