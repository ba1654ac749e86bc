"""Drop-in `sampler` package (reference: adelacvg/diff-vits `sampler/`): same module paths and public names
(`NoiseScheduleVP`, `model_wrapper`, `DPM_Solver`, `UniPC`, `interpolate_fn`, `expand_dims`). The multistep
DPM-Solver++ / UniPC loops run natively (one CUDA graph) when the model is a B200-native denoiser."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.append(_root)
