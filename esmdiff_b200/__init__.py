"""esmdiff_b200 -- B200-native implementation of ESMDiff's masked-diffusion sampling path
(``slm/sample_esmdiff.py --mode ddpm`` of lujiarui/esmdiff).

Host side: Python mirroring the reference's interfaces for this path
(``MaskedDiffusionLanguageModeling.ddpm_sample``, ``CustomizedESM3.forward``, ``TimestepEmbedder``,
``LogLinearNoise``, ``load_state_dict_from_lightning_ckpt``, the ``sample_esmdiff`` CLI).
Device side: hand-written sm_100a CUDA kernels (tcgen05 / TMEM / TMA) behind a C ABI
(``include/esmdiff_b200.h`` -> ``esmdiff_b200/lib/libesmdiff_b200.so``).
"""
from ._lib import EsmdiffError, build  # noqa: F401

__all__ = ["EsmdiffError", "build"]
