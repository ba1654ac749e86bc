"""Chained-GEMM kernels (csrc/chain_tc.cu) against the one-kernel-per-op program of the same transformer block.

The two programs round to bf16 at the same points and accumulate in the same K order, so every transformer block
output must agree to LayerNorm-summation-order noise (a last-bit difference in a mean can flip single bf16 roundings
downstream), far below the BF16 floor of the network. Shapes cover full tiles, ragged last tiles (L % 128 != 0 at
every level) and tiles shorter than one 32-row slab.
"""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(unet, B, T, Tp, fuse, precision="bf16"):
    from diffvits_b200.engine import DenoiserEngine
    old = os.environ.get("DV_FUSE")
    os.environ["DV_FUSE"] = "1" if fuse else "0"
    try:
        return DenoiserEngine(unet, B, T, Tp, "cuda:0", x_channels=100, precision=precision)
    finally:
        if old is None:
            os.environ.pop("DV_FUSE", None)
        else:
            os.environ["DV_FUSE"] = old


@pytest.fixture(scope="module")
def unet(seeded_unet):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return copy.deepcopy(seeded_unet).to("cuda")


@pytest.mark.parametrize("B,T,Tp", [(2, 256, 40), (3, 200, 282), (1, 1000, 64), (2, 40, 16)])
def test_fused_blocks_match_unfused(unet, B, T, Tp):
    import gpu_util as U
    from oracle import inputs
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(B, T, Tp, seed=11))
    t = torch.tensor([437.0], device="cuda")
    outs, named = [], []
    for fuse in (True, False):
        eng = _engine(unet, B, T, Tp, fuse)
        kinds = [o["kind"] for o in eng.op_table()]
        assert (5 in kinds) == fuse
        eng.prepare_prompt(ehs)
        eng.set_cond(cond)
        outs.append(eng.forward(x, t, x_only=True).clone())
        torch.cuda.synchronize()
        named.append({k: v.float().clone() for k, v in eng.named_tensors().items()})
    assert set(named[0]) == set(named[1])
    worst = 0.0
    for k in named[0]:
        r = U.rel_l2(named[0][k], named[1][k])
        worst = max(worst, r)
        print(f"  {k:40s} fused vs unfused rel {r:.3e}")
    # the first transformer block sees identical inputs: only LayerNorm summation order differs
    assert U.rel_l2(named[0]["down_blocks.0.attentions.0"], named[1]["down_blocks.0.attentions.0"]) < 2e-3
    assert worst < 2e-2                 # deeper blocks: rounding flips amplified to the BF16 floor (test_gpu_unet.py)
    assert torch.isfinite(outs[0]).all()
    assert U.rel_l2(outs[0], outs[1]) < 2e-2


def test_projection_layernorm_chain_matches_separate_kernels(unet, monkeypatch):
    """DV_FUSE_LN=1: proj_in / to_out + the LayerNorm that follows as a one-stage chain (CH_ROW with a global LayerNorm
    output) at the levels the full chain does not cover (C = 256, 384)."""
    import gpu_util as U
    from oracle import inputs
    B, T, Tp = 2, 200, 40
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(B, T, Tp, seed=13))
    t = torch.tensor([250.0], device="cuda")
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("DV_FUSE_LN", flag)
        eng = _engine(unet, B, T, Tp, False)
        tags = [o["tag"] for o in eng.op_table()]
        assert any(tg.endswith("+ln") for tg in tags) == (flag == "1")
        eng.prepare_prompt(ehs)
        eng.set_cond(cond)
        outs.append(eng.forward(x, t, x_only=True).clone())
        torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all()
    assert U.rel_l2(outs[0], outs[1]) < 2e-2


def test_fused_deterministic(unet):
    from oracle import inputs
    B, T, Tp = 2, 200, 40
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(B, T, Tp, seed=12))
    t = torch.tensor([100.0], device="cuda")
    eng = _engine(unet, B, T, Tp, True)
    eng.prepare_prompt(ehs)
    eng.set_cond(cond)
    a = eng.forward(x, t, x_only=True).clone()
    b = eng.forward(x, t, x_only=True).clone()
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,T,Tp,fuse,precision", [(2, 256, 40, True, "bf16"), (3, 203, 282, True, "bf16"), (5, 1000, 64, False, "bf16"),
                                                   (2, 40, 16, True, "bf16"), (40, 1000, 33, True, "bf16"), (3, 203, 282, True, "fp16"),
                                                   (3, 203, 282, True, "fp16s"), (2, 256, 40, True, "fp32x3"),
                                                   (5, 1000, 64, True, "fp32x3")])
def test_groupnorm_in_producer_epilogue_matches_standalone_pass(unet, monkeypatch, B, T, Tp, fuse, precision):
    """GroupNorm applied by the PRODUCING GEMM's CTAs (per-batch-item arrival counter, kernels.h GemmGnOut; default) against
    the stand-alone gn_apply launches (DV_GN_OUT=0): same stored bf16 input, same partial statistics folded in the same
    order, same coefficient arithmetic, so every block output must be bit-identical. Shapes: one and several row tiles
    per item, ragged last tiles, odd lengths, more tiles than CTA slots (B = 40 x 8 tiles at level 0 = 320 tiles on 296
    slots: the persistent tile loop goes round twice and an item straddles the rounds), and the one-kernel-per-op
    transformer program (GroupNorm fused into proj_out / conv2 at every level), and every precision mode (hi + lo planes of
    the split modes included). The plan's health flag (a bounded in-kernel wait that gave up) must stay clear."""
    from oracle import inputs
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(B, T, Tp, seed=21))
    t = torch.tensor([611.0], device="cuda")
    outs, named, n_gn = [], [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("DV_GN_OUT", flag)
        eng = _engine(unet, B, T, Tp, fuse, precision)
        ops = eng.op_table()
        n_gn.append(sum(1 for o in ops if o["kind"] == 2))
        assert any("+gn:" in o["tag"] for o in ops) == (flag == "1")
        eng.prepare_prompt(ehs)
        eng.set_cond(cond)
        for _ in range(3):                       # the item counters must return to zero after every launch
            out = eng.forward(x, t, x_only=True).clone()
        outs.append(out)
        torch.cuda.synchronize()
        assert int(eng.health_flag()) == 0
        named.append({k: v.float().clone() for k, v in eng.named_tensors().items()})
        eng.close()
    print(f"gn_apply launches per call: fused {n_gn[0]}, stand-alone {n_gn[1]}")
    assert n_gn[0] <= 16 < n_gn[1]
    for k in named[0]:
        assert torch.equal(named[0][k], named[1][k]), k
    assert torch.equal(outs[0], outs[1])


def test_fused_groupnorm_with_two_plans_running_concurrently(unet):
    """The item wait of the fused output GroupNorm assumes that the CTAs of one launch become co-resident. Two plans enqueued on
    two streams at the same time compete for the SMs: the waits must still complete (a straddling item only delays the CTAs of
    that item; everything else finishes and frees SMs), the health flags stay clear and the results equal the ones computed alone."""
    import diffvits_b200
    from oracle import inputs
    B, T, Tp = 16, 1000, 282
    betas = inputs.betas_linear()
    data = [tuple(t.cuda() for t in inputs.make_inputs(B, T, Tp, seed=40 + i)) for i in range(2)]
    nats = [diffvits_b200.NativeDenoiser(unet, d[1], d[2]) for d in data]
    alone = [diffvits_b200.sample_native(unet, d[0], d[1], d[2], sampler="unipc", steps=4, betas=betas, native=n, graph=False).clone()
             for d, n in zip(data, nats)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(3):
        outs = []
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for d, n, s in zip(data, nats, streams):
            with torch.cuda.stream(s):
                outs.append(diffvits_b200.sample_native(unet, d[0], d[1], d[2], sampler="unipc", steps=4, betas=betas, native=n,
                                                        graph=(rep > 0)))
        torch.cuda.synchronize()
        for o, a, n in zip(outs, alone, nats):
            n.engine.check_finite()                      # raises if a bounded wait gave up
            assert torch.equal(o, a)
