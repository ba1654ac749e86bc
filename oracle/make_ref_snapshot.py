"""ORACLE tooling (test infrastructure): snapshot the UNMODIFIED reference files of the hot path into `oracle/_ref/`.

    python -m oracle.make_ref_snapshot            # in the build container, where /root/reference exists

`/root/reference` does not exist on the GPU box. `oracle/_ref/` is git-ignored (the reference's sources never enter this
repository's history) but NOT gpurun-ignored, so the snapshot travels with the working tree the way the built `.so` does.
It holds exactly the files `import model3` loads plus the two sampler files and `config.json`,
byte for byte, with a SHA-256 manifest. Consumers (all test / measurement infrastructure, never the product path):

  * `bench.py --impl reference` and the `cpu_baseline` leg: the reference's own `Diffusion_Encoder` (per-step
    PromptEncoder included) + `sampler/*` on the host cores (`cpu_baseline.kind: "reference"`);
  * `bench.py`'s `gpu_reference` leg: the same modules eager on the B200 (fp32 / TF32 / autocast-bf16);
  * `tests/test_gpu_reference.py`: `model3.NaturalSpeech2.sample` over the drop-in classes vs over the reference classes.

`__graft_entry__.build()` refreshes the snapshot whenever /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("DIFFVITS_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")

# what `import model3` pulls in from the checkout (probe: sys.modules after the import) ...
FILES = [
    "model3.py", "model.py", "operations.py", "commons.py", "modules.py", "attentions.py", "dataset.py", "losses.py",
    "transforms.py", "utils.py", "parametrizations.py", "parametrize.py",
    "monotonic_align/__init__.py", "monotonic_align/core.py", "text/__init__.py", "text/symbols.py",
    "unet1d/__init__.py", "unet1d/activations.py", "unet1d/attention.py", "unet1d/attention_processor.py",
    "unet1d/dual_transformer_1d.py", "unet1d/embeddings.py", "unet1d/lora.py", "unet1d/outputs.py", "unet1d/resnet.py",
    "unet1d/transformer_1d.py", "unet1d/unet_1d_blocks.py", "unet1d/unet_1d_condition.py",
    # ... the samplers model3.sample imports lazily (model3.py:1128, 1162) and the model configuration
    "sampler/dpm_solver.py", "sampler/uni_pc.py", "config.json",
]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def available_source() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model3.py"))


def make(verbose: bool = True) -> str:
    if not available_source():
        raise SystemExit(f"reference checkout not found at {REF_SRC}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "adelacvg/diff-vits (unmodified files)", "files": manifest}, f, indent=1)
    if verbose:
        print(f"snapshot: {len(FILES)} files -> {DST}")
    return DST


if __name__ == "__main__":
    make()
    sys.exit(0)
