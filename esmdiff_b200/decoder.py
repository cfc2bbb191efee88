"""Batched VQ-VAE structure decoding + PDB writing: what the reference does after "Sampling token
time" (slm/sample_esmdiff.py:225-231 -> ``decode`` :41-61 -> ``esm3_model.decode(prot)`` ->
``raw_protein.to_pdb(tmp)`` per sample, then ``merge_pdbfiles``, eval_utils.py:437-492).

The reference decodes the samples ONE BY ONE (B = 1, a temporary PDB file each, then a merge);
with the token sampling at ~2 s for 100 samples that serial loop dominates its "Total time".  Here
all samples of a target go through the decoder in one batch on the same tcgen05 kernels as the
sampling network and the multi-MODEL PDB is written directly.

Only the structure half of ``ESM3.decode`` exists on this path (the sampler leaves every other
track at its default): backbone N / CA / C from ``StructureTokenDecoder`` +
``Dim6RotStructureHead``, the carbonyl O from ``ProteinChain.infer_oxygen``, pLDDT in the B-factor
column as ``ESMProtein.to_pdb`` writes it.  pTM / PAE (``pairwise_classification_head``) are not
written to a PDB and are not computed.  The esm package is not vendored in the reference and not
installed here: architecture restated from esm==3.0.4, parity unpinned.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .engine import DecoderDims, Engine
from .tokenization import STRUCTURE_BOS, STRUCTURE_EOS

ONE_TO_THREE = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY",
                "H": "HIS", "I": "ILE", "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER",
                "T": "THR", "W": "TRP", "Y": "TYR", "V": "VAL", "U": "SEC", "O": "PYL"}
ATOM_NAMES = ("N", "CA", "C", "O")


class StructureTokenDecoder:
    """``esm.models.vqvae.StructureTokenDecoder`` surface (``decode``, ``load_state_dict``) over a
    decoder context of the CUDA library.  State-dict keys are esm's (``embed.weight``,
    ``decoder_stack.blocks.{i}.*``, ``affine_output_projection.*``, ``plddt_head.*``)."""

    def __init__(self, d_model=1280, n_heads=20, n_layers=30, device=None, plddt_bins=50):
        self.dims = DecoderDims(d_model=d_model, n_heads=n_heads, n_layers=n_layers, plddt_bins=plddt_bins)
        self.engine = Engine(self.dims, device=device)

    @property
    def device(self):
        return self.engine.device

    def load_state_dict(self, state_dict, strict=True):
        self.engine.load_state_dict(state_dict, strict=strict)
        return self

    def eval(self):
        return self

    @torch.no_grad()
    def decode(self, structure_tokens: torch.Tensor, attention_mask=None, sequence_id=None) -> dict:
        """(B, T) int64 with BOS first and EOS last in every row (esm asserts the same).  Returns
        esm's keys; ``ptm`` / ``predicted_aligned_error`` are None (not on this path)."""
        assert attention_mask is None and sequence_id is None, "padding / multi-chain batches are not on this path"
        assert bool((structure_tokens[:, 0] == STRUCTURE_BOS).all()) and bool((structure_tokens[:, -1] == STRUCTURE_EOS).all()), \
            "structure tokens must start with BOS and end with EOS"
        assert int((structure_tokens < 0).sum()) == 0
        bb, o, plddt, _ = self.engine.decode_structure(structure_tokens)
        return {"tensor7_affine": None, "bb_pred": bb, "oxygen": o, "plddt": plddt, "ptm": None,
                "predicted_aligned_error": None}


def pdb_model_lines(sequence: str, bb: np.ndarray, o: np.ndarray, plddt: np.ndarray | None) -> list[str]:
    """ATOM records of one decoded chain (chain A, residues 1..L; N, CA, C, O per residue, atoms with
    NaN coordinates left out as biotite does for atom37 masks), then TER."""
    lines, serial = [], 1
    last = None
    for i, aa in enumerate(sequence):
        res = ONE_TO_THREE.get(aa, "UNK")
        b = float(plddt[i]) if plddt is not None else 0.0
        for a, name in enumerate(ATOM_NAMES):
            xyz = bb[i, a] if a < 3 else o[i]
            if not np.isfinite(xyz).all():
                continue
            lines.append(f"ATOM  {serial:5d}  {name:<3s} {res:>3s} A{i + 1:4d}    "
                         f"{xyz[0]:8.3f}{xyz[1]:8.3f}{xyz[2]:8.3f}{1.0:6.2f}{b:6.2f}          {name[0]:>2s}  ")
            last = (serial, res, i + 1)
            serial += 1
    if last is not None:
        lines.append(f"TER   {last[0] + 1:5d}      {last[1]:>3s} A{last[2]:4d}")
    return lines


@torch.no_grad()
def decode_to_pdb(decoder: StructureTokenDecoder, structure_tokens: torch.Tensor, sequence: str, save_to: Path,
                  max_tokens_per_batch: int = 1 << 17):
    """structure_tokens (N, L) WITHOUT BOS/EOS (what the sampler returns, sample_esmdiff.py:217-221)
    -> one multi-MODEL PDB at ``save_to`` in the layout of the reference's ``merge_pdbfiles``
    (MODEL n / ATOM ... / TER / ENDMDL per sample, END, lines padded to 80 columns).
    Returns (bb (N,L,3,3), plddt (N,L)) on the host."""
    N, L = structure_tokens.shape
    assert L == len(sequence), f"{L} structure tokens for a sequence of {len(sequence)} residues"
    tok = torch.cat([torch.full((N, 1), STRUCTURE_BOS, dtype=torch.int64), structure_tokens.cpu().to(torch.int64),
                     torch.full((N, 1), STRUCTURE_EOS, dtype=torch.int64)], dim=1)
    per = max(1, max_tokens_per_batch // (L + 2))
    bbs, os_, pls = [], [], []
    for i in range(0, N, per):                       # one batch unless N * (L + 2) is very large
        out = decoder.decode(tok[i:i + per])
        bbs.append(out["bb_pred"][:, 1:-1])
        os_.append(out["oxygen"][:, 1:-1])
        if out["plddt"] is not None:
            pls.append(out["plddt"][:, 1:-1])
    decoder.engine.synchronize()
    bb = torch.cat(bbs).cpu().numpy()
    ox = torch.cat(os_).cpu().numpy()
    pl = torch.cat(pls).cpu().numpy() if pls else None
    lines = []
    for n in range(N):
        lines.append(f"MODEL     {n + 1}")
        lines += [ln.strip() for ln in pdb_model_lines(sequence, bb[n], ox[n], pl[n] if pl is not None else None)]
        lines.append("ENDMDL")
    lines.append("END")
    save_to = Path(save_to)
    save_to.parent.mkdir(parents=True, exist_ok=True)
    save_to.write_text("\n".join(ln.ljust(80) for ln in lines) + "\n")
    return bb, pl


def load_decoder(path=None, device=None, seed: int = 0) -> StructureTokenDecoder:
    """The decoder of ``ESM3.get_structure_decoder()`` (``data/weights/esm3_structure_decoder_v0.pth``
    of the esm3_sm_open_v1 release).  ``path``: that state-dict file.  None -> random-init weights of
    the same architecture (the pretrained file cannot be fetched offline); the PDBs then hold
    geometry of an untrained decoder and are only good for throughput / plumbing."""
    from .synthetic import random_decoder_state_dict
    dec = StructureTokenDecoder(device=device)
    if path is not None:
        sd = torch.load(path, map_location="cpu", weights_only=False)
        sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    else:
        sd = random_decoder_state_dict(dec.dims, device=dec.device, seed=seed)
    dec.load_state_dict(sd)
    return dec
