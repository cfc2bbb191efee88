"""Import the reference's OWN sampler code verbatim (ORACLE tooling -- build container only).

``/root/reference/slm/models/model.py`` and ``slm/utils/noise_utils.py`` are imported from
where they lie, unmodified, through stub modules for the packages that are absent here
(SURVEY.md 8c): ``esm.utils.constants.esm3`` (11 constants), ``lightning.LightningModule``,
``torchmetrics``; ``slm.models.net`` is reduced to ``TimestepEmbedder`` (net.py:486-522) and
``slm.models.utils`` to ``cross_entropy`` (models/utils.py:197-201), both exec'd from the
reference source text at import time.  Nothing is copied into this repository.

/root/reference does not exist on the GPU box: only ``oracle/make_golden.py`` and the
container-only tests (skipped when the tree is absent) use this module.
"""
from __future__ import annotations

import ast
import importlib.util
import sys
import types
from pathlib import Path

import torch
from torch import nn

REF_ROOT = Path("/root/reference")


def available() -> bool:
    return (REF_ROOT / "slm/models/model.py").exists()


def _extract(src_path: Path, names: set[str]) -> str:
    """Source text of the named top-level defs/classes of a reference file."""
    text = src_path.read_text()
    tree = ast.parse(text)
    lines = text.splitlines()
    out = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            out.append("\n".join(lines[node.lineno - 1:node.end_lineno]))
    assert len(out) == len(names), f"missing {names} in {src_path}"
    return "\n\n".join(out)


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "slm.models.model" in sys.modules:
        return
    consts = dict(SEQUENCE_BOS_TOKEN=0, SEQUENCE_PAD_TOKEN=1, SEQUENCE_EOS_TOKEN=2,
                  SEQUENCE_CHAINBREAK_TOKEN=31, SEQUENCE_MASK_TOKEN=32,
                  VQVAE_CODEBOOK_SIZE=4096, STRUCTURE_MASK_TOKEN=4096, STRUCTURE_EOS_TOKEN=4097,
                  STRUCTURE_BOS_TOKEN=4098, STRUCTURE_PAD_TOKEN=4099,
                  STRUCTURE_CHAINBREAK_TOKEN=4100)
    esm3 = _module("esm.utils.constants.esm3", **consts)
    _module("esm", __path__=[])
    _module("esm.utils", __path__=[])
    _module("esm.utils.constants", __path__=[], esm3=esm3)

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def save_hyperparameters(self, *a, **k):
            pass

    _module("lightning", LightningModule=LightningModule)

    class _Metric:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            pass

        def reset(self):
            pass

        def compute(self):
            return 0.0

    _module("torchmetrics", MeanMetric=_Metric, MinMetric=_Metric)

    _module("slm", __path__=[str(REF_ROOT / "slm")])
    _module("slm.utils", __path__=[str(REF_ROOT / "slm/utils")])      # skip its hydra-laden __init__
    _module("slm.models", __path__=[str(REF_ROOT / "slm/models")])

    import math
    import torch.nn.functional as F
    net = _module("slm.models.net")
    net.__dict__.update(math=math, torch=torch, nn=nn)
    exec(_extract(REF_ROOT / "slm/models/net.py", {"TimestepEmbedder"}), net.__dict__)
    mu = _module("slm.models.utils")
    mu.__dict__.update(F=F, torch=torch)
    exec(_extract(REF_ROOT / "slm/models/utils.py", {"cross_entropy"}), mu.__dict__)

    for mod, rel in (("slm.utils.noise_utils", "slm/utils/noise_utils.py"),
                     ("slm.models.model", "slm/models/model.py")):
        spec = importlib.util.spec_from_file_location(mod, REF_ROOT / rel)
        m = importlib.util.module_from_spec(spec)
        sys.modules[mod] = m
        spec.loader.exec_module(m)
        if mod == "slm.utils.noise_utils":
            sys.modules["slm.utils"].noise_utils = m


def load():
    """Returns (model_module, noise_utils_module, TimestepEmbedder) of the reference."""
    assert available(), "/root/reference is not present (GPU box?)"
    _install_stubs()
    return (sys.modules["slm.models.model"], sys.modules["slm.utils.noise_utils"],
            sys.modules["slm.models.net"].TimestepEmbedder)


def build_reference_sampler(net: nn.Module, sigma_embedder: nn.Module):
    """The reference's MaskedDiffusionLanguageModeling wired as configs/experiment/mdlm.yaml
    :26-58 wires it, around any ``net``."""
    model_mod, noise_mod, _ = load()
    cls = model_mod.MaskedDiffusionLanguageModeling
    # LanguageModeling.__init__ signature: read it rather than guess
    import inspect
    base_params = inspect.signature(model_mod.LanguageModeling.__init__).parameters
    kw = {}
    if "net" in base_params:
        kw["net"] = net
    for name in ("optimizer", "scheduler"):
        if name in base_params:
            kw[name] = None
    if "compile" in base_params:
        kw["compile"] = False
    m = cls(noise_schedule=noise_mod.LogLinearNoise(), sigma_embedder=sigma_embedder,
            time_conditioning=True, change_of_variables=False, importance_sampling=False,
            condition_dropout=0.0, condition_mask_rate=0.0, sequence_prediction=False,
            T=0, sampling_eps=1e-3, noise_removal=True, structure_only=False,
            coupled_condition_mask=False, **kw)
    m.noise_removal = True      # checkpoint_utils.py:71
    return m.eval()
