"""Checkpoint -> model for the sampling CLI: the boundary of reference
``slm/utils/checkpoint_utils.py:41-74`` (``load_state_dict_from_lightning_ckpt``).

The reference reads the ``model:`` block of ``.hydra/config.yaml`` (or
``configs/experiment/mdlm.yaml``) with OmegaConf and builds it with ``hydra.utils.instantiate``;
neither package is needed for that: the block is plain YAML with ``_target_`` keys, resolved here
against this package's classes.  The state dict is the DeepSpeed ``['module']`` dict; its key
names are the weight ABI.
"""
from __future__ import annotations

import re
from pathlib import Path

import torch
import yaml

from . import model as _model
from . import net as _net
from . import noise_utils as _noise

# `_target_` strings the reference's configs use -> classes of this package
TARGETS = {
    "slm.models.model.MaskedDiffusionLanguageModeling": _model.MaskedDiffusionLanguageModeling,
    "slm.models.net.CustomizedESM3": _net.CustomizedESM3,
    "slm.models.net.TimestepEmbedder": _net.TimestepEmbedder,
    "slm.utils.noise_utils.LogLinearNoise": _noise.LogLinearNoise,
    "slm.utils.noise_utils.CosineNoise": _noise.CosineNoise,
}

# the model: block of configs/experiment/mdlm.yaml:26-58, used when no config file is found
DEFAULT_MODEL_CFG = {
    "_target_": "slm.models.model.MaskedDiffusionLanguageModeling",
    "compile": False, "optimizer": {"lr": 1e-5}, "scheduler": None,
    "noise_schedule": {"_target_": "slm.utils.noise_utils.LogLinearNoise"},
    "T": 0, "noise_removal": True, "sampling_eps": 1e-3, "time_conditioning": True,
    "change_of_variables": False, "importance_sampling": False, "sequence_prediction": False,
    "condition_dropout": 0.0, "condition_mask_rate": 0.0, "coupled_condition_mask": False,
    "structure_only": False,
    "sigma_embedder": {"_target_": "slm.models.net.TimestepEmbedder", "hidden_size": 1536},
    "net": {"_target_": "slm.models.net.CustomizedESM3", "pretrained": True,
            "n_structure_heads": 4101, "n_sequence_heads": 0},
}

_NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


def _coerce(v):
    """PyYAML (YAML 1.1) reads ``1e-5`` as a string; OmegaConf reads a float."""
    if isinstance(v, str) and _NUM.match(v):
        return float(v)
    return v


def instantiate(node, **overrides):
    """Minimal ``hydra.utils.instantiate`` for nested ``_target_`` dicts."""
    if isinstance(node, dict):
        if "_target_" in node:
            target = node["_target_"]
            if target not in TARGETS:
                raise ValueError(f"unsupported _target_ on the ddpm path: {target}")
            kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
            kwargs.update(overrides)
            return TARGETS[target](**kwargs)
        return {k: instantiate(v) for k, v in node.items()}
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    return _coerce(node)


def load_model_cfg(ckpt_path: Path) -> dict:
    if ckpt_path.is_dir():
        cfg_path = ckpt_path.parent.parent.parent / ".hydra/config.yaml"
    else:
        cfg_path = ckpt_path.parent.parent / ".hydra/config.yaml"
    candidates = [cfg_path, Path("configs/experiment/mdlm.yaml")]
    for p in candidates:
        if p.exists():
            cfg = yaml.safe_load(p.read_text())
            print(f"Loaded experiment config: {p}...")
            return cfg["model"]
    print(f"Config file not found: {cfg_path}. Use default config.")
    return dict(DEFAULT_MODEL_CFG)


def build_model(model_cfg: dict | None = None, device=None, **net_overrides):
    cfg = dict(model_cfg or DEFAULT_MODEL_CFG)
    net_cfg = dict(cfg.pop("net"))
    net_cfg.pop("_target_", None)
    net = _net.CustomizedESM3(**{k: _coerce(v) for k, v in net_cfg.items()}, device=device,
                              time_conditioning=bool(cfg.get("time_conditioning", False)),
                              **net_overrides)
    return instantiate(cfg, net=net)


def load_state_dict_from_lightning_ckpt(ckpt_path, device="cuda"):
    print(f"Loading ESMDiff ckpt from {ckpt_path}")
    ckpt_path = Path(ckpt_path)
    assert ckpt_path.exists(), f"Checkpoint not found: {ckpt_path}"
    assert ckpt_path.suffix in [".ckpt", ".pt"], f"Unsupported ckpt format: {ckpt_path}"
    model_cfg = load_model_cfg(ckpt_path)
    if ckpt_path.is_dir():      # deepspeed checkpoint directory
        ckpt_path = ckpt_path / "checkpoint/mp_rank_00_model_states.pt"
    dev_index = torch.device(device).index if str(device) != "cuda" else None
    model = build_model(model_cfg, device=dev_index)
    print("Sucessfully instantiated model ...")
    if ckpt_path.suffix == ".pt":
        all_params = torch.load(ckpt_path, map_location="cpu", weights_only=False)["module"]
        model.load_state_dict(all_params)
    else:
        raise ValueError(f"Unsupported ckpt format: {ckpt_path}")
    print(f"Sucessfully loaded model from {ckpt_path}...")
    model.noise_removal = True      # necessary when decoding from tokens (checkpoint_utils.py:71)
    model.to(device)
    model.eval()
    return model
