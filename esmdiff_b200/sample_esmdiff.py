"""``python -m esmdiff_b200.sample_esmdiff`` -- the reference's sampling CLI
(slm/sample_esmdiff.py:236-294) for ``--mode ddpm``, same flags and defaults.

    --input DIR --ckpt F --output DIR --mode {gibbs,ddpm} --num_steps N --num_samples N --mask_ids a,b,c

Output layout as the reference: ``OUT/step{N}_eps{eps}_N{num}_{time}/{stem}.pdb`` (skipped when it
exists, :155-160).  The step after "Sampling token time" -- VQ-VAE structure decoding and PDB
writing (:225-231) -- needs the pretrained ESM3 structure decoder from the ``esm`` package, which
is outside this path: when ``esm`` is importable it is used exactly as the
reference uses it; otherwise the sampled structure tokens are written next to where the PDB
would go (``{stem}.structure_tokens.pt``) and the decode step is reported as skipped.
``--mode gibbs`` (the esm SDK's own sampler) is not part of this path and is refused.
"""
from __future__ import annotations

import argparse
from pathlib import Path
from time import strftime, time

import torch

from .checkpoint_utils import load_state_dict_from_lightning_ckpt
from .sampling import sample_structure_tokens
from .tokenization import sequence_from_pdb, tokenize_sequence


def _esm_available():
    try:
        import esm  # noqa: F401
        return True
    except Exception:
        return False


def _decode_with_esm(structure_tokens, sequence_tokens, save_paths):
    """sample_esmdiff.py:41-61 through the esm SDK (only when installed)."""
    from esm.models.esm3 import ESM3
    from esm.sdk.api import ESMProteinTensor
    from .tokenization import (SEQUENCE_BOS, SEQUENCE_EOS, STRUCTURE_BOS, STRUCTURE_EOS)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    esm3 = ESM3.from_pretrained("esm3_sm_open_v1").to(dev)
    seq = torch.cat([torch.LongTensor([SEQUENCE_BOS]), sequence_tokens.cpu(), torch.LongTensor([SEQUENCE_EOS])])
    for st, path in zip(structure_tokens, save_paths):
        s = torch.cat([torch.LongTensor([STRUCTURE_BOS]), st.cpu(), torch.LongTensor([STRUCTURE_EOS])])
        esm3.decode(ESMProteinTensor(sequence=seq, structure=s).to(dev)).to_pdb(path)


def merge_pdbfiles(pdb_files, save_to: Path):
    """Ordered merge into one multi-MODEL file (behaviour of eval_utils.py:437-492 for
    single-model inputs, which is all the CLI produces)."""
    lines, n = [], 0
    for f in pdb_files:
        n += 1
        lines.append(f"MODEL     {n}")
        lines += [ln.strip() for ln in Path(f).read_text().splitlines() if ln.startswith(("TER", "ATOM"))]
        lines.append("ENDMDL")
    lines.append("END")
    save_to.parent.mkdir(parents=True, exist_ok=True)
    save_to.write_text("\n".join(ln.ljust(80) for ln in lines) + "\n")


@torch.no_grad()
def ddpm_sample_by_esm(sequence, pl_model, output_dir: Path, sample_basename: str, num_samples=5,
                       num_steps=10, eps=1e-5, mask_ids=None, structure_tokens=None, sample_max_t=1.0):
    str_time = strftime("%Y%m%d-%H%M%S")
    output_dir = output_dir / f"step{num_steps}_eps{eps}_N{num_samples}_{str_time}"
    save_to = output_dir / f"{sample_basename}.pdb"
    print(f"Results will save to {save_to}")
    if save_to.exists():
        print(f"Skip existing {save_to}")
        return None
    output_dir.mkdir(parents=True, exist_ok=True)
    if mask_ids is not None:
        assert structure_tokens is not None, \
            "inpainting needs structure tokens of the known residues (VQ-VAE encoder output)"
        seq = list(sequence)
        for idx in mask_ids:
            assert 0 <= idx < len(seq), f"Invalid mask index {idx} for sequence of length {len(seq)}"
            seq[idx] = "_"
        sequence = "".join(seq)
    seq_tokens = tokenize_sequence(sequence)
    start_t = time()
    tokens, _ = sample_structure_tokens(pl_model, seq_tokens, num_samples, num_steps, eps=eps,
                                        structure_tokens=structure_tokens, mask_ids=mask_ids,
                                        sample_max_t=sample_max_t)
    tokens = tokens.cpu()
    if _esm_available():
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            paths = [Path(tmp) / f"{sample_basename}.{i}.pdb" for i in range(len(tokens))]
            _decode_with_esm(tokens, seq_tokens[1:-1], paths)
            merge_pdbfiles(paths, save_to)
    else:
        tok_path = output_dir / f"{sample_basename}.structure_tokens.pt"
        torch.save({"sequence": sequence, "sequence_tokens": seq_tokens, "structure_tokens": tokens}, tok_path)
        print(f"esm package not installed: structure decode skipped, tokens saved to {tok_path}")
    print(f"Total time: {time() - start_t:.2f}s")
    return tokens


def get_argparser():
    p = argparse.ArgumentParser(description="Evaluate the ensemble of protein structures.")
    p.add_argument("--input", type=str, default="data/targets/bpti", help="Path to the data directory.")
    p.add_argument("--ckpt", type=str, default=None, help="Path to the model checkpoint.")
    p.add_argument("--output", type=str, default="output/inference_esmdiff")
    p.add_argument("--mode", type=str, default="gibbs", choices=["gibbs", "ddpm"])
    p.add_argument("--num_steps", type=int, default=25, help="Number of denoising steps.")
    p.add_argument("--num_samples", type=int, default=10, help="Number of samples to generate.")
    p.add_argument("--mask_ids", type=str, default=None, help="Comma-separated list of masked indices.")
    p.add_argument("--prior_tokens", type=str, default=None,
                   help="(extension) .pt with 'structure_tokens' (L+2,) for --mask_ids inpainting when "
                        "the esm VQ-VAE encoder is not installed")
    return p


def main(argv=None):
    args = get_argparser().parse_args(argv)
    if args.mode != "ddpm":
        raise SystemExit("esmdiff_b200 implements --mode ddpm only (gibbs is the esm SDK's sampler, "
                         "outside this path)")
    assert args.ckpt is not None, "--mode ddpm needs --ckpt (sample_esmdiff.py:252-258)"
    model = load_state_dict_from_lightning_ckpt(args.ckpt, device="cuda")
    data_path = Path(args.input)
    assert data_path.is_dir(), f"Invalid directory {data_path} (Currently we only support pdb files in a folder as input)."
    print(f">>> Sampling mode = {args.mode} ...")
    output_dir = Path(args.output)
    output_dir.mkdir(parents=True, exist_ok=True)
    prior = None
    if args.prior_tokens:
        prior = torch.load(args.prior_tokens, weights_only=False)["structure_tokens"].to(torch.int64)
    for p in [q for q in data_path.iterdir() if q.suffix == ".pdb"]:
        sequence = sequence_from_pdb(p)
        mask_ids = [int(i) for i in args.mask_ids.split(",")] if args.mask_ids is not None else None
        ddpm_sample_by_esm(sequence, model, output_dir, p.stem, num_samples=args.num_samples,
                           num_steps=args.num_steps, mask_ids=mask_ids, structure_tokens=prior)


if __name__ == "__main__":
    main()
