"""The slice of the ``esm`` SDK's operator surface the sampling path touches, over the CUDA library:
``ESMProtein`` (``from_pdb``, ``to_pdb``, ``sequence``, ``coordinates``), ``ESMProteinTensor`` and an ``ESM3``-shaped
object with ``encode`` / ``decode`` / ``to`` -- so that the reference's own helpers run unchanged on top of it:

    slm/sample_esmdiff.py:41-61    decode(structure_tokens, sequence_tokens, esm3_model, save_to)
                                     -> ESMProteinTensor(sequence=, structure=).to(device); esm3_model.decode(prot).to_pdb(save_to)
    slm/sample_esmdiff.py:278-284  prot = ESMProtein.from_pdb(p); prot.sequence; prot.coordinates
    slm/models/utils.py:136-137    prot = ESMProtein(sequence=, coordinates=); gt_tokens = model.encode(prot)
                                     -> gt_tokens.sequence, gt_tokens.structure

The reference imports these from ``esm.sdk.api`` / ``esm.models.esm3`` (esm==3.0.4, not vendored, not installed).
Only the sequence and structure tracks exist here -- the ones the path reads; ``encode`` runs the VQ-VAE structure
encoder (esmdiff_b200/encoder.py), ``decode`` the structure decoder (esmdiff_b200/decoder.py), both on the GPU.
The batched product path (``decode_to_pdb``: all samples of a target in one call) does not go through this
per-sample surface; it exists for drop-in use of the reference's helper functions and for tests.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from pathlib import Path

import torch

from .tokenization import (AA_TO_ID, SEQUENCE_BOS, SEQUENCE_EOS, SEQUENCE_VOCAB, STRUCTURE_BOS, STRUCTURE_EOS,
                           tokenize_sequence)


@dataclass
class ESMProtein:
    """``esm.sdk.api.ESMProtein``: sequence (one letter per residue, '_' = masked) and atom37 coordinates (L, 37, 3)."""
    sequence: str | None = None
    coordinates: torch.Tensor | None = None
    plddt: torch.Tensor | None = None

    def __len__(self):
        if self.sequence is not None:
            return len(self.sequence)
        if self.coordinates is not None:
            return self.coordinates.size(0)
        raise ValueError("No track to determine length from.")

    @classmethod
    def from_pdb(cls, path, chain_id: str = "detect") -> "ESMProtein":
        from .encoder import coordinates_from_pdb
        seq, coords = coordinates_from_pdb(Path(path), None if chain_id == "detect" else chain_id)
        return cls(sequence=seq, coordinates=coords)

    def to_pdb(self, path) -> None:
        """Backbone N / CA / C / O of chain A with pLDDT in the B-factor column (what ``ESMProtein.to_pdb`` writes for
        a decoded structure)."""
        from .decoder import pdb_model_lines
        assert self.coordinates is not None and self.sequence is not None
        c = self.coordinates.detach().float().cpu().numpy()
        pl = None if self.plddt is None else self.plddt.detach().float().cpu().numpy()
        lines = pdb_model_lines(self.sequence.replace("_", "X"), c[:, :3], c[:, 4], pl)
        Path(path).write_text("\n".join(ln.ljust(80) for ln in lines + ["END"]) + "\n")


@dataclass
class ESMProteinTensor:
    """``esm.sdk.api.ESMProteinTensor``: the tokenised tracks (BOS / EOS included)."""
    sequence: torch.Tensor | None = None
    structure: torch.Tensor | None = None
    coordinates: torch.Tensor | None = None

    def __len__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                return v.size(0)
        raise ValueError("No track to determine length from.")

    def to(self, device) -> "ESMProteinTensor":
        return ESMProteinTensor(*[None if getattr(self, f.name) is None else getattr(self, f.name).to(device)
                                  for f in fields(self)])


class ESM3:
    """``esm.models.esm3.ESM3`` as the sampling path uses it: an ``esm3_model`` with ``encode`` / ``decode`` / ``to``.
    ``structure_encoder`` / ``structure_decoder``: esmdiff_b200.encoder.StructureTokenEncoder / decoder.StructureTokenDecoder
    (``get_structure_encoder()`` / ``get_structure_decoder()`` of esm)."""

    def __init__(self, structure_encoder=None, structure_decoder=None):
        self._structure_encoder = structure_encoder
        self._structure_decoder = structure_decoder

    def to(self, device):
        return self                       # weights already live on the contexts' device

    def get_structure_encoder(self):
        assert self._structure_encoder is not None, "no structure encoder loaded (--encoder_ckpt)"
        return self._structure_encoder

    def get_structure_decoder(self):
        assert self._structure_decoder is not None, "no structure decoder loaded (--decoder_ckpt)"
        return self._structure_decoder

    @torch.no_grad()
    def encode(self, input: ESMProtein) -> ESMProteinTensor:
        """esm ``ESM3.encode``: tokenize_sequence (BOS / EOS, '_' -> mask id 32) and, given coordinates,
        tokenize_structure through the VQ-VAE encoder; coordinates padded with inf rows for BOS / EOS."""
        from .encoder import tokenize_structure
        seq = None if input.sequence is None else tokenize_sequence(input.sequence)
        structure = coords = None
        if input.coordinates is not None:
            structure = tokenize_structure(input.coordinates, self.get_structure_encoder())
            c = input.coordinates
            coords = torch.full((c.size(0) + 2,) + tuple(c.shape[1:]), float("inf"), dtype=c.dtype)
            coords[1:-1] = c
        return ESMProteinTensor(sequence=seq, structure=structure, coordinates=coords)

    @torch.no_grad()
    def decode(self, input: ESMProteinTensor) -> ESMProtein:
        """esm ``ESM3.decode`` for the two tracks of the path: sequence ids -> letters, structure tokens ->
        backbone coordinates (atom37 with N, CA, C, O filled) + pLDDT."""
        sequence = None
        if input.sequence is not None:
            ids = input.sequence.cpu().tolist()
            assert ids[0] == SEQUENCE_BOS and ids[-1] == SEQUENCE_EOS, "sequence tokens must carry BOS / EOS"
            sequence = "".join("_" if i == AA_TO_ID["_"] else SEQUENCE_VOCAB[i] for i in ids[1:-1])
        coordinates = plddt = None
        if input.structure is not None:
            st = input.structure
            assert int(st[0]) == STRUCTURE_BOS and int(st[-1]) == STRUCTURE_EOS, "structure tokens must carry BOS / EOS"
            out = self.get_structure_decoder().decode(st[None])
            self.get_structure_decoder().engine.synchronize()
            bb, ox = out["bb_pred"][0, 1:-1].cpu(), out["oxygen"][0, 1:-1].cpu()
            coordinates = torch.full((bb.size(0), 37, 3), float("nan"))
            coordinates[:, :3] = bb
            coordinates[:, 4] = ox
            plddt = None if out["plddt"] is None else out["plddt"][0, 1:-1].cpu()
        return ESMProtein(sequence=sequence, coordinates=coordinates, plddt=plddt)


@torch.no_grad()
def decode(structure_tokens, sequence_tokens, esm3_model: ESM3, save_to=None):
    """reference slm/sample_esmdiff.py:41-61, per sample (the CLI itself decodes all samples in one batch)."""
    assert len(structure_tokens) == len(sequence_tokens), f"{len(structure_tokens)} != {len(sequence_tokens)}"
    sequence_tokens = torch.cat([torch.LongTensor([SEQUENCE_BOS]), sequence_tokens.cpu(), torch.LongTensor([SEQUENCE_EOS])])
    structure_tokens = torch.cat([torch.LongTensor([STRUCTURE_BOS]), structure_tokens.cpu(), torch.LongTensor([STRUCTURE_EOS])])
    prot = ESMProteinTensor(sequence=sequence_tokens, structure=structure_tokens)
    raw_protein = esm3_model.decode(prot)
    if save_to is not None:
        raw_protein.to_pdb(save_to)
    return raw_protein
