"""``python -m esmdiff_b200.sample_esmdiff`` -- the reference's sampling CLI
(slm/sample_esmdiff.py:236-294) for ``--mode ddpm``, same flags and defaults.

    --input DIR --ckpt F --output DIR --mode {gibbs,ddpm} --num_steps N --num_samples N --mask_ids a,b,c

Output layout as the reference: ``OUT/step{N}_eps{eps}_N{num}_{time}/{stem}.pdb`` (skipped when it
exists, :155-160), a multi-MODEL file in ``merge_pdbfiles``' layout.  The step after "Sampling
token time" -- VQ-VAE structure decoding and PDB writing (:225-231) -- runs BATCHED on the same
CUDA kernels (esmdiff_b200/decoder.py) when decoder weights are given:
    --decoder_ckpt PATH     state dict of esm's ``StructureTokenDecoder`` (the file the esm package
                            fetches as data/weights/esm3_structure_decoder_v0.pth)
    --decoder_ckpt random   random-init weights of that architecture (plumbing / throughput runs)
Without it: when the ``esm`` package is importable it is used exactly as the reference uses it
(serial B=1 decodes); otherwise the sampled structure tokens are written next to where the PDB would
go (``{stem}.structure_tokens.pt``) and the decode step is reported as skipped.
``--mask_ids a,b,c`` (inpainting, :197-201) needs the structure tokens of the known residues: ``--encoder_ckpt PATH|random``
runs the VQ-VAE structure ENCODER on the PDB's backbone coordinates as the reference does (esmdiff_b200/encoder.py),
``--prior_tokens F`` takes them from a file.
``--mode gibbs`` (the reference's default: the esm SDK's entropy-ordered iterative sampler driving the same
network, sample_esmdiff.py:66-130) runs through esmdiff_b200/gibbs.py on one GPU; it needs ``--ckpt`` (the
pretrained ESM3 weights the reference falls back to cannot be fetched offline).  With ``--mask_ids`` (and
``--encoder_ckpt``) the known residues' coordinates condition every forward through block 0's geometric attention
and their VQ-VAE codes form the structure prompt.
Multi-GPU: ``torchrun --nproc-per-node N -m esmdiff_b200.sample_esmdiff ...`` shards the samples of
every target over the ranks; rank 0 decodes and writes.
"""
from __future__ import annotations

import argparse
from pathlib import Path
from time import strftime, time

import torch

from .checkpoint_utils import load_state_dict_from_lightning_ckpt
from . import distributed as D
from .sampling import sample_structure_tokens_sharded
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


def timer(func):
    """eval_utils.py:24-34: prints ``Elapsed time (name): x.xx sec`` when the call returned something.  The reference's
    wrapper also re-packs the result as ``(*result, elapsed)``; its callers ignore the value, so this one returns the
    function's own result."""
    import functools

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        start_time = time()
        result = func(*args, **kwargs)
        if result is None:
            return None
        if kwargs.get("rank", 0) == 0:
            print(f"Elapsed time ({func.__name__}): {time() - start_time:.2f} sec")
        return result
    return wrapper


def merge_pdbfiles(pdb_files, save_to: Path):
    """Ordered merge into one multi-MODEL file (behaviour of eval_utils.py:437-492 for
    single-model inputs, which is all the CLI produces -- including the closing ENDMDL the reference
    appends after the last model's own: the file ends ENDMDL / ENDMDL / END; pinned by
    tests/golden/merged_models.pdb, the reference function's own output)."""
    lines, n = [], 0
    for f in pdb_files:
        n += 1
        lines.append(f"MODEL     {n}")
        lines += [ln.strip() for ln in Path(f).read_text().splitlines() if ln.startswith(("TER", "ATOM"))]
        lines.append("ENDMDL")
    lines.append("ENDMDL")
    lines.append("END")
    save_to.parent.mkdir(parents=True, exist_ok=True)
    save_to.write_text("\n".join(ln.ljust(80) for ln in lines) + "\n")


@timer
@torch.no_grad()
def ddpm_sample_by_esm(sequence, pl_model, output_dir: Path, sample_basename: str, num_samples=5,
                       num_steps=10, eps=1e-5, mask_ids=None, structure_tokens=None, sample_max_t=1.0,
                       rank=0, world=1, seed=None, decoder=None, coordinates=None, esm3_model=None):
    """reference sample_esmdiff.py:137-233.  ``rank`` / ``world`` (torchrun, one process per GPU):
    every rank samples its share of ``num_samples``; rank 0 alone decodes and writes.
    ``coordinates`` (L, A, 3) + ``esm3_model`` (here: the VQ-VAE structure encoder, esmdiff_b200/encoder.py): the
    reference's inpainting front end, ``protseq_to_data(sequence, esm3_model, encode_only=True, coordinates=,
    mask_ids=)`` (:166-174); ``structure_tokens`` (L + 2,) given directly takes its place."""
    str_time = strftime("%Y%m%d-%H%M%S")
    output_dir = output_dir / f"step{num_steps}_eps{eps}_N{num_samples}_{str_time}"
    save_to = output_dir / f"{sample_basename}.pdb"
    skip = save_to.exists()
    if world > 1:                      # one decision for all ranks (rank 0's clock and file system view)
        box = [skip]
        torch.distributed.broadcast_object_list(box, src=0)
        skip = box[0]
    if rank == 0:
        print(f"Results will save to {save_to}")
    if skip:
        if rank == 0:
            print(f"Skip existing {save_to}")
        return None
    if rank == 0:
        output_dir.mkdir(parents=True, exist_ok=True)
    if mask_ids is not None and structure_tokens is None:
        assert coordinates is not None and esm3_model is not None, \
            "inpainting needs the structure tokens of the known residues: coordinates + the VQ-VAE encoder " \
            "(--encoder_ckpt) or --prior_tokens"
        from .encoder import protseq_to_data
        structure_tokens = protseq_to_data(sequence, esm3_model, encode_only=True, coordinates=coordinates,
                                           mask_ids=mask_ids)["structure_tokens"]
    if mask_ids is not None:
        seq = list(sequence)
        for idx in mask_ids:
            assert 0 <= idx < len(seq), f"Invalid mask index {idx} for sequence of length {len(seq)}"
            seq[idx] = "_"
        sequence = "".join(seq)
    seq_tokens = tokenize_sequence(sequence)
    start_t = time()
    tokens, _ = sample_structure_tokens_sharded(pl_model, seq_tokens, num_samples, num_steps, rank=rank, world=world,
                                                seed=seed, eps=eps, structure_tokens=structure_tokens,
                                                mask_ids=mask_ids, sample_max_t=sample_max_t, verbose=rank == 0)
    tokens = tokens.cpu()
    if rank != 0:
        return tokens
    if decoder is not None:
        from .decoder import decode_to_pdb
        decode_to_pdb(decoder, tokens, "".join(sequence), save_to)
    elif _esm_available():
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
    p.add_argument("--seed", type=int, default=None,
                   help="(extension) seed torch before sampling; under torchrun rank r uses seed + its first "
                        "sample index.  Default: unseeded, as the reference")
    p.add_argument("--decoder_ckpt", type=str, default=None,
                   help="(extension) esm StructureTokenDecoder state dict for the built-in batched decode, or "
                        "'random' for random-init weights of that architecture")
    p.add_argument("--encoder_ckpt", type=str, default=None,
                   help="(extension) esm StructureTokenEncoder state dict (data/weights/esm3_structure_encoder_v0.pth) "
                        "for --mask_ids inpainting: the known residues' structure tokens come from the PDB's "
                        "coordinates as in the reference; 'random' = random-init weights of that architecture")
    p.add_argument("--prior_tokens", type=str, default=None,
                   help="(extension) .pt with 'structure_tokens' (L+2,) for --mask_ids inpainting instead of "
                        "--encoder_ckpt")
    return p


def _main_gibbs(args):
    """sample_esmdiff.py:257-294 with sample_fn = minibatch_gibbs_by_esm(esm3_model=model.net); one GPU."""
    from .gibbs import minibatch_gibbs_by_esm
    model = load_state_dict_from_lightning_ckpt(args.ckpt, device="cuda")
    data_path = Path(args.input)
    assert data_path.is_dir(), f"Invalid directory {data_path} (Currently we only support pdb files in a folder as input)."
    print(f">>> Sampling mode = {args.mode} ...")
    output_dir = Path(args.output)
    output_dir.mkdir(parents=True, exist_ok=True)
    decoder = None
    if args.decoder_ckpt:
        from .decoder import load_decoder
        decoder = load_decoder(None if args.decoder_ckpt == "random" else args.decoder_ckpt)
    encoder = None
    if args.mask_ids is not None:
        assert args.encoder_ckpt, "--mode gibbs --mask_ids conditions on the PDB's coordinates: give --encoder_ckpt PATH|random"
        from .encoder import load_encoder
        encoder = load_encoder(None if args.encoder_ckpt == "random" else args.encoder_ckpt)
    for p in sorted(q for q in data_path.iterdir() if q.suffix == ".pdb"):
        coordinates = mask_ids = None
        if args.mask_ids is not None:
            from .encoder import coordinates_from_pdb
            mask_ids = [int(i) for i in args.mask_ids.split(",")]           # 0-based index (:281)
            sequence, coordinates = coordinates_from_pdb(p)                 # prot.sequence, prot.coordinates (:278-283)
        else:
            sequence = sequence_from_pdb(p)
        minibatch_gibbs_by_esm(sequence, model.net, output_dir, p.stem, num_samples=args.num_samples,
                               num_steps=args.num_steps, coordinates=coordinates, mask_ids=mask_ids, decoder=decoder,
                               seed=args.seed, structure_encoder=encoder)


def main(argv=None):
    args = get_argparser().parse_args(argv)
    assert args.ckpt is not None, ("--ckpt is required: the pretrained ESM3 weights the reference loads without it "
                                   "(sample_esmdiff.py:36, :252-255) cannot be fetched offline")
    if args.mode == "gibbs":
        return _main_gibbs(args)
    # `torchrun --nproc-per-node N -m esmdiff_b200.sample_esmdiff ...`: one process per GPU, the
    # samples of every target sharded over the ranks; plain `python -m` = one GPU
    rank, world, local = D.init_from_env()
    device = f"cuda:{local}" if world > 1 else "cuda"
    model = load_state_dict_from_lightning_ckpt(args.ckpt, device=device)
    data_path = Path(args.input)
    assert data_path.is_dir(), f"Invalid directory {data_path} (Currently we only support pdb files in a folder as input)."
    if rank == 0:
        print(f">>> Sampling mode = {args.mode} ..." + (f" ({world} GPUs)" if world > 1 else ""))
    output_dir = Path(args.output)
    if rank == 0:
        output_dir.mkdir(parents=True, exist_ok=True)
    decoder = None
    if args.decoder_ckpt and rank == 0:              # rank 0 alone decodes and writes
        from .decoder import load_decoder
        decoder = load_decoder(None if args.decoder_ckpt == "random" else args.decoder_ckpt,
                               device=local if world > 1 else None)
    prior = None
    if args.prior_tokens:
        prior = torch.load(args.prior_tokens, weights_only=False)["structure_tokens"].to(torch.int64)
    encoder = None
    if args.encoder_ckpt and args.mask_ids is not None and prior is None:
        from .encoder import load_encoder
        encoder = load_encoder(None if args.encoder_ckpt == "random" else args.encoder_ckpt,
                               device=local if world > 1 else None)
    for p in sorted(q for q in data_path.iterdir() if q.suffix == ".pdb"):
        coordinates = None
        if args.mask_ids is not None:
            mask_ids = [int(i) for i in args.mask_ids.split(",")]           # 0-based index (:281)
            if encoder is not None:
                from .encoder import coordinates_from_pdb
                sequence, coordinates = coordinates_from_pdb(p)             # prot.sequence, prot.coordinates (:278-283)
            else:
                sequence = sequence_from_pdb(p)
        else:
            mask_ids = None
            sequence = sequence_from_pdb(p)
        ddpm_sample_by_esm(sequence, model, output_dir, p.stem, num_samples=args.num_samples,
                           num_steps=args.num_steps, mask_ids=mask_ids, structure_tokens=prior,
                           rank=rank, world=world, seed=args.seed, decoder=decoder, coordinates=coordinates,
                           esm3_model=encoder)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
