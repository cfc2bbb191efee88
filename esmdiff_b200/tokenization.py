"""Sequence tokenisation and a minimal PDB sequence reader for the CLI front end.

The reference gets both from the ``esm`` SDK (``ESMProtein.from_pdb(p).sequence``,
``ESM3.encode``; sample_esmdiff.py:278, models/utils.py:136-137), which is not vendored.  The id
map below is pinned by the reference's own fixtures (``data/dummy_train_data/*.pth``:
``sequence`` <-> ``sequence_tokens``; frozen in tests/golden/tokenizer_pins.json).
"""
from __future__ import annotations

from pathlib import Path

import torch

SEQUENCE_BOS, SEQUENCE_PAD, SEQUENCE_EOS, SEQUENCE_MASK = 0, 1, 2, 32
STRUCTURE_MASK, STRUCTURE_EOS, STRUCTURE_BOS, STRUCTURE_PAD, STRUCTURE_CHAINBREAK = 4096, 4097, 4098, 4099, 4100

# ESM3 sequence vocabulary (esm.utils.constants.esm3.SEQUENCE_VOCAB)
SEQUENCE_VOCAB = ["<cls>", "<pad>", "<eos>", "<unk>", "L", "A", "G", "V", "S", "E", "R", "T", "I", "D",
                  "P", "K", "Q", "N", "F", "Y", "M", "H", "W", "C", "X", "B", "U", "Z", "O", ".", "-",
                  "|", "<mask>"]
AA_TO_ID = {a: i for i, a in enumerate(SEQUENCE_VOCAB)}
AA_TO_ID["_"] = SEQUENCE_MASK          # ESMProtein writes masked residues as '_'

THREE_TO_ONE = {"ALA": "A", "ARG": "R", "ASN": "N", "ASP": "D", "CYS": "C", "GLN": "Q", "GLU": "E",
                "GLY": "G", "HIS": "H", "ILE": "I", "LEU": "L", "LYS": "K", "MET": "M", "PHE": "F",
                "PRO": "P", "SER": "S", "THR": "T", "TRP": "W", "TYR": "Y", "VAL": "V", "SEC": "U",
                "PYL": "O", "MSE": "M"}


def tokenize_sequence(sequence: str) -> torch.Tensor:
    """BOS + residues + EOS, int64 (L+2,)."""
    ids = [SEQUENCE_BOS] + [AA_TO_ID.get(ch, AA_TO_ID["X"]) for ch in sequence] + [SEQUENCE_EOS]
    return torch.tensor(ids, dtype=torch.int64)


def sequence_from_pdb(path: Path, chain: str | None = None) -> str:
    """One-letter sequence of the first (or named) chain of the first model, from ATOM records."""
    seq, seen, first_chain = [], set(), None
    for line in Path(path).read_text().splitlines():
        if line.startswith("ENDMDL"):
            break
        if not line.startswith(("ATOM", "HETATM")) or len(line) < 27:
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
        if key in seen:
            continue
        seen.add(key)
        seq.append(THREE_TO_ONE.get(resname, "X"))
    return "".join(seq)


def synthetic_sequence_tokens(L: int, seed: int = 0) -> torch.Tensor:
    """BOS + L ids uniform in {4..23} + EOS (SURVEY.md 8d synthetic inputs)."""
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(4, 24, (L,), generator=g)
    return torch.cat([torch.tensor([SEQUENCE_BOS]), body, torch.tensor([SEQUENCE_EOS])]).to(torch.int64)
