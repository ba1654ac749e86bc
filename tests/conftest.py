import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dropin():
    """The drop-in unet1d / sampler packages, ahead of anything else on sys.path."""
    import diffvits_b200
    diffvits_b200.install_dropin()
    import sampler.dpm_solver as dpm
    import sampler.uni_pc as uni
    import unet1d.unet_1d_condition as u
    assert "dropin" in u.__file__
    return dict(unet=u, dpm=dpm, uni=uni)


@pytest.fixture(scope="session")
def seeded_unet(dropin):
    """Drop-in UNet with the reference's seeded default init (seed 1234); verified against the golden checksums."""
    import json

    import torch
    from oracle.ref_import import UNET_KW
    torch.manual_seed(1234)
    m = dropin["unet"].UNet1DConditionModel(**UNET_KW).eval()
    with open(os.path.join(GOLD, "weights_seed1234.json")) as f:
        g = json.load(f)
    sd = m.state_dict()
    assert len(sd) == g["n_tensors"]
    for k, (s, a) in g["sums"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), k
    return m
