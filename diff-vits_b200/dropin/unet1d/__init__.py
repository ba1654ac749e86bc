"""Drop-in `unet1d` package: same module paths, class names, constructor kwargs and state_dict layout as
adelacvg/diff-vits `unet1d/`, with the forward pass executed by libdiffvits_b200 (sm_100a CUDA kernels).

Put this directory's parent (`diff-vits_b200/dropin`) ahead of the reference checkout on PYTHONPATH (or call
`diffvits_b200.install_dropin()` before `import model3`); `model3.py` / `tts_infer.py` then resolve
`unet1d.unet_1d_condition.UNet1DConditionModel` and `sampler.*` to these classes unchanged.
"""
import os as _os
import sys as _sys

# make `import diffvits_b200` work when only the dropin directory was put on PYTHONPATH
_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.append(_root)
