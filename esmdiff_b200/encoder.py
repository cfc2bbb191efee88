"""VQ-VAE structure ENCODER front end: what the reference does before sampling when coordinates matter
(``--mask_ids`` inpainting): ``ESMProtein.from_pdb(p).coordinates`` (slm/sample_esmdiff.py:278-284),
``protseq_to_data`` (slm/models/utils.py:105-146: masked residues -> '_' / ``coordinates[idx] = inf``,
``model.encode(ESMProtein(sequence, coordinates))``) and ``pdb_to_data`` (:99-102).

The reference gets all of it from the ``esm`` SDK (``ESM3.encode`` -> ``tokenize_structure`` ->
``StructureTokenEncoder.encode``), which is not vendored; here the encoder runs on the CUDA library
(csrc/encoder.cu: fp32 -- the result is a code index) behind esm's surface (``encode``, ``load_state_dict`` with
esm's state-dict keys).  Architecture restated from esm==3.0.4, parity unpinned.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import torch

from . import _lib
from ._lib import EncoderCfg, EsmdiffError
from .engine import _ptr, _stream
from .tokenization import (STRUCTURE_BOS, STRUCTURE_EOS, THREE_TO_ONE, tokenize_sequence)

# atom37 order of esm.utils.residue_constants.atom_types (ESMProtein.coordinates is (L, 37, 3))
ATOM37 = ("N", "CA", "C", "CB", "O", "CG", "CG1", "CG2", "OG", "OG1", "SG", "CD", "CD1", "CD2", "ND1", "ND2", "OD1",
          "OD2", "SD", "CE", "CE1", "CE2", "CE3", "NE", "NE1", "NE2", "OE1", "OE2", "CH2", "NH1", "NH2", "OH", "CZ",
          "CZ2", "CZ3", "NZ", "OXT")
ATOM37_INDEX = {a: i for i, a in enumerate(ATOM37)}


@dataclass
class EncoderDims:
    """esm ``StructureTokenEncoder(d_model=1024, n_heads=1, v_heads=128, n_layers=2, d_out=128, n_codes=4096)``
    (ESM3_structure_encoder_v0), knn = 16, RelativePositionEmbedding(32, d_model)."""
    d_model: int = 1024
    v_heads: int = 128
    n_layers: int = 2
    d_out: int = 128
    n_codes: int = 4096
    knn: int = 16
    rel_bins: int = 32

    @property
    def ffn_hidden(self) -> int:
        return int(((8.0 / 3.0 * self.d_model) + 255) // 256 * 256)


class StructureTokenEncoder:
    """``esm.models.vqvae.StructureTokenEncoder`` surface over an encoder context of the CUDA library."""

    def __init__(self, dims: EncoderDims | None = None, device: int | None = None):
        self.dims = dims or EncoderDims()
        self.L = _lib.lib()
        if not torch.cuda.is_available():
            raise EsmdiffError("esmdiff_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        d = self.dims
        cfg = EncoderCfg(d.d_model, d.v_heads, d.n_layers, d.ffn_hidden, d.d_out, d.n_codes, d.knn, d.rel_bins)
        h = C.c_void_p()
        if self.L.esmdiff_encoder_create(C.byref(cfg), self.device_index, C.byref(h)) != 0:
            raise EsmdiffError(self.L.esmdiff_encoder_last_error(None).decode())
        self.h = h
        self.codebook = None

    def _check(self, rc: int):
        if rc != 0:
            raise EsmdiffError(self.L.esmdiff_encoder_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.esmdiff_encoder_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self):
        return self

    def load_state_dict(self, state_dict, strict: bool = True):
        for k, v in state_dict.items():
            t = v.detach().float().contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
            self._check(self.L.esmdiff_encoder_set_weight(self.h, k.encode(), _ptr(t), int(t.is_cuda), 0, shape,
                                                          max(t.dim(), 1)))
        self._check(self.L.esmdiff_encoder_finalize(self.h))
        self.codebook = state_dict["codebook.embeddings"].detach().float().to(self.device)
        return self

    @torch.no_grad()
    def encode(self, coords: torch.Tensor, attention_mask=None, sequence_id=None, residue_index=None,
               return_aux: bool = False):
        """coords (B, L, >=3, 3): N, CA, C first (atom3 / atom14 / atom37), NaN or inf = unknown.
        Returns esm's ``(z_q (B, L, d_out), min_encoding_indices (B, L) int64)``."""
        assert attention_mask is None and sequence_id is None, "padding / multi-chain batches are not on this path"
        assert coords.dim() == 4 and coords.size(-1) == 3 and coords.size(-2) >= 3, "need N, CA, C"
        B, L = coords.shape[:2]
        c = coords[..., :3, :].to(self.device, torch.float32).contiguous()
        ri = None if residue_index is None else residue_index.to(self.device, torch.int64).contiguous()
        E = min(self.dims.knn, L)
        codes = torch.empty(B, L, dtype=torch.int64, device=self.device)
        z = torch.empty(B, L, self.dims.d_out, dtype=torch.float32, device=self.device)
        edges = torch.empty(B, L, E, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.L.esmdiff_encode_structure(self.h, _ptr(c), _ptr(ri), B, L, _ptr(codes), _ptr(z),
                                                        _ptr(edges), _stream()))
        if return_aux:
            return {"z": z, "codes": codes, "edges": edges}
        return self.codebook[codes], codes


def normalize_coordinates(coords: torch.Tensor) -> torch.Tensor:
    """esm ``normalize_coordinates`` (``ProteinChain.to_structure_encoder_inputs``): the chain expressed in the
    frame of its average backbone (Gram-Schmidt of mean C -> CA, CA -> N; origin mean CA).  (L, A, 3) -> same."""
    bb = coords[:, :3, :].double()
    ok = torch.isfinite(bb).all(-1).all(-1)
    if not bool(ok.any()):
        return coords
    avg = bb[ok].mean(0)
    e0 = avg[1] - avg[2]
    e0 = e0 / (e0.pow(2).sum() + 1e-12).sqrt()
    e1 = avg[0] - avg[1]
    e1 = e1 - e0 * (e0 * e1).sum()
    e1 = e1 / (e1.pow(2).sum() + 1e-12).sqrt()
    rot = torch.stack([e0, e1, torch.linalg.cross(e0, e1)], dim=-1)            # columns e0 e1 e2
    return ((coords.double() - avg[1]) @ rot).to(coords.dtype)


@torch.no_grad()
def tokenize_structure(coordinates: torch.Tensor, structure_encoder: StructureTokenEncoder) -> torch.Tensor:
    """esm ``tokenize_structure``: (L, A, 3) coordinates of one chain -> int64 (L + 2,) BOS, codes, EOS (host)."""
    L = coordinates.size(0)
    c = normalize_coordinates(coordinates.float())
    _, codes = structure_encoder.encode(c[None], residue_index=torch.arange(1, L + 1)[None])
    out = torch.empty(L + 2, dtype=torch.int64)
    out[0], out[-1] = STRUCTURE_BOS, STRUCTURE_EOS
    out[1:-1] = codes[0].cpu()
    return out


@torch.no_grad()
def protseq_to_data(sequence: str, model: Optional[StructureTokenEncoder], coordinates: torch.Tensor | None = None,
                    encode_only: bool = False, mask_ids: Optional[list] = None, filled_ids: Optional[list] = None,
                    total_size: Optional[int] = None):
    """reference slm/models/utils.py:105-146 for ``encode_only=True`` (the only mode the sampling CLI uses).
    ``model``: what stands in for ``esm3_model`` here -- the structure encoder (``model.encode`` of the reference
    tokenises the sequence and, given coordinates, runs this encoder)."""
    assert encode_only, "only encode_only=True is on the sampling path (sample_esmdiff.py:166-174)"
    if coordinates is not None:
        coordinates = coordinates.clone()
    if mask_ids is not None:
        sequence = list(sequence)
        for idx in mask_ids:
            assert 0 <= idx < len(sequence), f"Invalid mask index {idx} for sequence of length {len(sequence)}"
            sequence[idx] = "_"
            coordinates[idx] = float("Inf")
        sequence = "".join(sequence)
    elif filled_ids is not None:
        assert total_size is not None, "total_size must be provided when fill_ids is not None"
        assert all(0 <= idx < total_size for idx in filled_ids), f"Invalid fill index {filled_ids} for sequence of length {total_size}"
        _seq = ["_"] * total_size
        _coord = coordinates.new_ones(total_size, coordinates.size(1), 3) * float("Inf")
        for idx in filled_ids:
            _seq[idx] = sequence[idx]
            _coord[idx] = coordinates[idx]
        sequence = "".join(_seq)
        coordinates = _coord
    structure_tokens = None
    if coordinates is not None:
        assert model is not None, "coordinates need the VQ-VAE structure encoder (--encoder_ckpt)"
        structure_tokens = tokenize_structure(coordinates, model)
    return {"sequence_tokens": tokenize_sequence(sequence), "structure_tokens": structure_tokens,
            "sequence": sequence, "coordinates": coordinates}


def coordinates_from_pdb(path: Path, chain: str | None = None):
    """(sequence, coordinates (L, 37, 3) float32 atom37 with NaN for absent atoms) of the first (or named) chain of
    the first model: what ``ESMProtein.from_pdb(path)`` exposes as ``.sequence`` / ``.coordinates``.  Alternate
    locations: the first one listed wins."""
    seq, rows, index, first_chain = [], [], {}, None
    for line in Path(path).read_text().splitlines():
        if line.startswith("ENDMDL"):
            break
        if not line.startswith(("ATOM", "HETATM")) or len(line) < 54:
            continue
        ch = line[21]
        if chain is not None and ch != chain:
            continue
        if first_chain is None:
            first_chain = ch
        if chain is None and ch != first_chain:
            continue
        resname = line[17:20].strip()
        if line.startswith("HETATM") and resname not in THREE_TO_ONE:
            continue
        key = (ch, line[22:27])
        if key not in index:
            index[key] = len(seq)
            seq.append(THREE_TO_ONE.get(resname, "X"))
            rows.append(torch.full((37, 3), float("nan")))
        name = line[12:16].strip()
        if resname == "MSE" and name == "SE":
            name = "SD"
        a = ATOM37_INDEX.get(name)
        if a is not None and torch.isnan(rows[index[key]][a, 0]):
            rows[index[key]][a] = torch.tensor([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    coords = torch.stack(rows) if rows else torch.zeros(0, 37, 3)
    return "".join(seq), coords


@torch.no_grad()
def pdb_to_data(pdb_file, model: Optional[StructureTokenEncoder] = None, **kwargs):
    """reference slm/models/utils.py:99-102."""
    sequence, coordinates = coordinates_from_pdb(pdb_file)
    return protseq_to_data(sequence=sequence, model=model, coordinates=coordinates, **kwargs)


def load_encoder(path=None, device=None, seed: int = 0) -> StructureTokenEncoder:
    """The encoder of ``ESM3.get_structure_encoder()`` (``data/weights/esm3_structure_encoder_v0.pth`` of the
    esm3_sm_open_v1 release).  ``path``: that state-dict file.  None -> random-init weights of the same
    architecture (the pretrained file cannot be fetched offline): plumbing / throughput only."""
    from .synthetic import random_encoder_state_dict
    enc = StructureTokenEncoder(device=device)
    if path is not None:
        sd = torch.load(path, map_location="cpu", weights_only=False)
        sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    else:
        sd = random_encoder_state_dict(enc.dims, device=enc.device, seed=seed)
    enc.load_state_dict(sd)
    return enc
