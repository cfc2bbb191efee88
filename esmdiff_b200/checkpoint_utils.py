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
    """The ``model:`` block of the run's ``.hydra/config.yaml``, two levels above the checkpoint --
    the same directory for a single-file checkpoint and for a DeepSpeed checkpoint DIRECTORY
    (reference checkpoint_utils.py:45-49: four parents of ``dir/checkpoint/mp_rank_00_model_states.pt``
    = two parents of ``dir``) -- else ``configs/experiment/mdlm.yaml``, else the built-in copy of it."""
    cfg_path = ckpt_path.parent.parent / ".hydra/config.yaml"
    candidates = [cfg_path, Path("configs/experiment/mdlm.yaml")]
    for p in candidates:
        if p.exists():
            cfg = yaml.safe_load(p.read_text())
            print(f"Loaded experiment config: {p}...")
            return cfg["model"]
    print(f"Config file not found: {cfg_path}. Use default config.")
    return dict(DEFAULT_MODEL_CFG)


# Keys of the model: block that only matter for training.  A run's composed .hydra/config.yaml
# carries configs/model/default.yaml underneath the experiment: ``optimizer`` / ``scheduler`` are
# ``_partial_`` torch / transformers factories there, and ``net.config`` is the T5 baseline's
# ``transformers.T5Config`` block, which ``CustomizedESM3.__init__(..., *args, **kwargs)`` swallows
# unused (net.py:323-334).  None of them is instantiated here.
TRAINING_ONLY_KEYS = ("optimizer", "scheduler", "compile")
NET_KEYS = ("d_model", "n_heads", "v_heads", "n_layers", "pretrained", "n_structure_heads", "n_sequence_heads")


def split_model_cfg(model_cfg: dict | None) -> tuple[dict, dict]:
    """(model kwargs incl. nested ``_target_`` nodes, CustomizedESM3 kwargs) from a ``model:`` block."""
    cfg = dict(model_cfg or DEFAULT_MODEL_CFG)
    target = cfg.get("_target_", "slm.models.model.MaskedDiffusionLanguageModeling")
    if target != "slm.models.model.MaskedDiffusionLanguageModeling":
        raise ValueError(f"the ddpm path samples MaskedDiffusionLanguageModeling checkpoints, not {target}")
    for k in TRAINING_ONLY_KEYS:
        cfg.pop(k, None)
    net_cfg = dict(cfg.pop("net", None) or {})
    nt = net_cfg.pop("_target_", "slm.models.net.CustomizedESM3")
    if nt != "slm.models.net.CustomizedESM3":
        raise ValueError(f"unsupported net on the ddpm path: {nt}")
    net_kwargs = {k: _coerce(v) for k, v in net_cfg.items() if k in NET_KEYS}
    return cfg, net_kwargs


def build_model(model_cfg: dict | None = None, device=None, **net_overrides):
    cfg, net_kwargs = split_model_cfg(model_cfg)
    net_kwargs.update(net_overrides)
    net = _net.CustomizedESM3(**net_kwargs, device=device, time_conditioning=bool(cfg.get("time_conditioning", False)))
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
