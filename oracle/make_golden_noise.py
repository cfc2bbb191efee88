"""Golden values of every noise schedule of the reference (slm/utils/noise_utils.py:122-213, imported verbatim through
``oracle/ref_loader.py``): total / rate on a time grid and the importance-sampling transformation where one exists.
TEST INFRASTRUCTURE; run in the build container:   python -m oracle.make_golden_noise  ->  tests/golden/noise_schedules.npz"""
from pathlib import Path

import numpy as np
import torch

from . import ref_loader


def main():
    _, nu, _ = ref_loader.load()
    t = torch.linspace(0.0, 1.0, 41)[:, None]
    ti = torch.linspace(0.001, 0.999, 33)
    out = {"t": t.numpy(), "t_importance": ti.numpy()}
    for name, kw in (("LogLinearNoise", {}), ("CosineNoise", {}), ("CosineSqrNoise", {}), ("Linear", {"sigma_min": 0.01, "sigma_max": 8.0}),
                     ("GeometricNoise", {"sigma_min": 1e-3, "sigma_max": 2.0})):
        n = getattr(nu, name)(**kw)
        total, rate = n(t)
        out[f"{name}_total"] = total.numpy()
        out[f"{name}_rate"] = (rate * torch.ones_like(t)).numpy()
        if hasattr(n, "importance_sampling_transformation"):
            out[f"{name}_importance"] = n.importance_sampling_transformation(ti).numpy()
    path = Path(__file__).resolve().parent.parent / "tests" / "golden" / "noise_schedules.npz"
    np.savez_compressed(path, **out)
    print(path, sorted(out))


if __name__ == "__main__":
    main()
