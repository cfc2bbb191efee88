"""Shared by the CPU and GPU suites: regenerate the seeded sampler cases of
tests/golden/sampler_seeded.npz (inputs come from stored seeds, expectations from the file)."""
import numpy as np
import pytest
import torch

MASK = 4096


def _seeded_case(seed, B, T, frac, scale):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, T, 4101, generator=g) * scale
    u = torch.rand(B, T, 4101, generator=g)
    x = torch.randint(0, 4096, (B, T), generator=g)
    x = torch.where(torch.rand(B, T, generator=g) < frac, torch.full_like(x, MASK), x)
    return logits, u, x


def seeded_cases(golden_dir):
    g = np.load(golden_dir / "sampler_seeded.npz")
    for i in range(int(g["n"])):
        c = {k[: -len(f"_{i}")]: g[k] for k in g.files if k.endswith(f"_{i}")}
        logits, u, x = _seeded_case(int(c["seed"]), int(c["B"]), int(c["T"]), float(c["frac"]), float(c["scale"]))
        if abs(float(logits.double().sum()) - float(c["logits_sum"])) > 1e-6 * max(1.0, abs(float(c["logits_sum"]))):
            pytest.skip("torch CPU generator stream differs from the one the fixtures were made with")
        assert np.array_equal(x.numpy(), c["x_t"])
        yield c, logits, u, x
