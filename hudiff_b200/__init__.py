"""hudiff_b200 -- B200-native (sm_100a) implementation of HuDiff's denoising sampling path.

Scope: the `AntiTFNet` / `NanoAntiTFNet` denoiser forward and the
order-agnostic sampling loop of `antibody_scripts/sample.py` and
`nanobody_scripts/nanosample.py`, behind the reference's Python entry points.
All device arithmetic lives in `libhudiff_b200.so` (hand-written CUDA, C-ABI in
`include/hudiff.h`); this package is the thin host mirror of the reference API.
"""
from .config import AB_RELEASE, NB_RELEASE, EasyDict, install_easydict_shim, release_config
from .modules import AntiTFNetB200, NanoAntiTFNetB200, model_selected
from .tokenizer import Tokenizer

__all__ = [
    "AB_RELEASE", "NB_RELEASE", "EasyDict", "install_easydict_shim", "release_config",
    "AntiTFNetB200", "NanoAntiTFNetB200", "model_selected", "Tokenizer", "seed_all",
]
__version__ = "0.1.0"


def seed_all(seed: int) -> None:
    """utils/misc.py:27-31 -- seeds torch (CPU + CUDA default generators), numpy, random."""
    import random

    import numpy as np
    import torch

    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
