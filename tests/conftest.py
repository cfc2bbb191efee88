import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def engine():
    """One full-dims engine (no weights) for the single-kernel tests."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from esmdiff_b200.engine import Dims, Engine
    eng = Engine(Dims())
    yield eng
    eng.close()


TINY = dict(d_model=256, n_heads=4, v_heads=8, n_layers=2)


@pytest.fixture(scope="session")
def tiny_pair():
    """(oracle net, oracle sigma embedder, engine) with the same seeded random weights, tiny dims."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from esmdiff_b200.engine import Dims, Engine
    from oracle import esm3_ref
    net, emb = esm3_ref.build_reference_model(esm3_ref.Esm3Dims(**TINY), seed=0)
    eng = Engine(Dims(**TINY))
    eng.load_state_dict(esm3_ref.full_state_dict(net, emb))
    yield net, emb, eng
    eng.close()
