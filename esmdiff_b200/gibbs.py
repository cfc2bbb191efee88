"""``--mode gibbs``: the reference's ``minibatch_gibbs_by_esm`` (slm/sample_esmdiff.py:66-130) on the
B200 path.  The reference hands the network to the esm SDK's ``iterative_sampling_raw`` with
``GenerationConfig(track="structure", num_steps, temperature, top_p)`` per sample; here the same
loop -- forward without time conditioning, top-p / temperature sampling of every masked position,
entropy-ordered unmasking on a cosine schedule -- runs device resident through
``esmdiff_gibbs_sample`` (csrc/gibbs.cuh) for ALL samples of a target in one batch, followed by the
batched structure decode.  esm==3.0.4 is not vendored in the reference: the sampler semantics are
restated (parity unpinned), the call-site contract (arguments, defaults, output
directory name, chunk list, skip-if-exists) is the reference's.

Inpainting (``--mask_ids`` in gibbs mode, sample_esmdiff.py:92-98): the masked residues lose their letter ('_') and
their coordinates (inf); esm then conditions on what is left -- the VQ-VAE encoder's codes of the known residues as the
structure prompt (esmdiff_b200/encoder.py) and the known backbone frames through block 0's geometric attention
(``esmdiff_set_structure_coords``), live at every step.  What esm 3.0.4 does with the codes of the frameless
residues is not pinned by anything in the reference tree; here they are MASK -- the positions the sampler fills.
"""
from __future__ import annotations

import math
from pathlib import Path
from time import strftime, time

import torch

from .tokenization import STRUCTURE_BOS, STRUCTURE_EOS, STRUCTURE_MASK, tokenize_sequence

N_MAX_RESIDUE_SQUARE = 200 * 200 * 105      # sample_esmdiff.py:76


def unmask_schedule(num_steps: int, total_to_sample: int) -> list[int]:
    """Positions revealed per step: esm's cosine schedule
    (``_get_iterative_sampling_mask_for_prompt_and_step``): after 0-based step t,
    ``int(cos(pi/2 (t+1)/N) * total + 0.1)`` positions stay masked (0 after the last step); N is capped at
    the number of masked positions (iterative_sampling_tokens).  fp32 tensor arithmetic as in esm."""
    if total_to_sample > 0:
        num_steps = min(num_steps, total_to_sample)
    ks, still = [], total_to_sample
    for t in range(num_steps):
        perc = torch.cos(torch.tensor((t + 1) / num_steps) * math.pi * 0.5)
        after = int((perc * torch.tensor(total_to_sample) + 0.1).int()) if t + 1 < num_steps else 0
        k = max(still - after, 0)
        ks.append(k)
        still -= k
    return ks


def gibbs_chunk_sizes(L: int, num_samples: int, n_max_residue_square: int = N_MAX_RESIDUE_SQUARE) -> list[int]:
    """The reference's batch list (sample_esmdiff.py:104-113; L = residues, no BOS/EOS)."""
    target = L * L * num_samples
    n_batch = target // n_max_residue_square
    batch_size = n_max_residue_square // int(L * L)
    bsz = [batch_size] * n_batch
    if target % n_max_residue_square > 0:
        bsz.append(num_samples - sum(bsz))
    assert sum(bsz) == num_samples, f"{sum(bsz)} != {num_samples}"
    return bsz


@torch.no_grad()
def gibbs_sample_structure_tokens(net, sequence_tokens_singleton: torch.Tensor, num_samples: int, num_steps: int,
                                  temperature: float = 1.4, top_p: float = 0.9, prior: torch.Tensor | None = None,
                                  seed: int = 0, rng: str = "philox", chunks: list[int] | None = None,
                                  structure_coords: torch.Tensor | None = None):
    """(tokens int64 (num_samples, L) without BOS/EOS, seconds).  ``net``: the CUDA network
    (``CustomizedESM3``; its engine runs the loop).  rng "philox": one device-resident call per chunk;
    "torch": Exp(1) draws from torch's generator, one ``exponential_`` per step like torch.multinomial.
    ``structure_coords`` (T, >=3, 3), BOS / EOS rows included (inf): the prompt's backbone, fed to every forward."""
    eng = net.engine
    T = sequence_tokens_singleton.size(0)
    if prior is None:
        prior = torch.full((T,), STRUCTURE_MASK, dtype=torch.int64)
        prior[0], prior[-1] = STRUCTURE_BOS, STRUCTURE_EOS
    total = int((prior[1:-1] == STRUCTURE_MASK).sum())
    ks = unmask_schedule(num_steps, total)
    if chunks is None:
        from .sampling import chunk_sizes_b200
        chunks = gibbs_chunk_sizes(T - 2, num_samples) if rng == "torch" else chunk_sizes_b200(T, num_samples)
    start_t = time()
    outs, done = [], 0
    for bs in chunks:
        seq = sequence_tokens_singleton[None, :].repeat(bs, 1)
        pr = prior[None, :].repeat(bs, 1)
        if structure_coords is not None:
            eng.set_structure_coords(structure_coords[None].expand(bs, *structure_coords.shape))
        if rng == "philox":
            outs.append(eng.gibbs_sample(seq, pr, ks, temperature, top_p, seed=seed + done))
        else:
            x = pr.to(eng.device).contiguous()
            seqd = seq.to(eng.device)
            for t, k in enumerate(ks):
                logits, _ = eng.forward(seqd, x, aux=None)
                noise = torch.empty_like(logits).exponential_()
                eng.gibbs_step(x, logits, noise, temperature, top_p, k)
            outs.append(x)
        done += bs
    if structure_coords is not None:
        eng.set_structure_coords(None)
    tokens = torch.cat(outs, dim=0)[:, 1:-1]
    torch.cuda.synchronize(tokens.device)
    eng.synchronize()
    return tokens, time() - start_t


def _timer(func):
    from .sample_esmdiff import timer              # lazy: sample_esmdiff imports this module's functions lazily too
    return timer(func)


@_timer
@torch.no_grad()
def minibatch_gibbs_by_esm(protseq, esm3_model, output_dir: Path, sample_basename: str, num_samples: int = 10,
                           num_steps: int = 16, temperature: float = 1.4, top_p: float = 0.9,
                           n_max_residue_square: int = N_MAX_RESIDUE_SQUARE, coordinates=None, mask_ids=None,
                           decoder=None, seed: int | None = None, structure_encoder=None):
    """reference sample_esmdiff.py:66-130.  ``esm3_model`` = the CUDA network (``model.net``);
    ``structure_encoder``: the VQ-VAE encoder ``esm3_model.encode`` would run on ``coordinates`` (inpainting)."""
    str_time = strftime("%Y%m%d-%H%M%S")
    output_dir = output_dir / f"T{temperature}_step{num_steps}_topp{top_p}_N{num_samples}_{str_time}"
    save_to = output_dir / f"{sample_basename}.pdb"
    print(f"Results will save to {save_to}")
    if save_to.exists():
        print(f"Skip existing {save_to}")
        return None
    prior = coords = None
    if mask_ids is not None:
        print(f"Masking {len(mask_ids)} residues and inpainting...")
        assert coordinates is not None, "Need to provide coordinates for masking"
        assert structure_encoder is not None, "gibbs inpainting needs the VQ-VAE structure encoder (--encoder_ckpt)"
        from .encoder import tokenize_structure
        protseq = list(protseq)
        coordinates = coordinates.clone()
        for idx in mask_ids:
            assert 0 <= idx < len(protseq), f"Invalid mask index {idx} for sequence of length {len(protseq)}"
            protseq[idx] = "_"
            coordinates[idx] = float("Inf")
        protseq = "".join(protseq)
    if coordinates is not None:
        assert structure_encoder is not None, "coordinates need the VQ-VAE structure encoder (--encoder_ckpt)"
        from .encoder import tokenize_structure
        prior = tokenize_structure(coordinates, structure_encoder)
        known = torch.isfinite(coordinates[:, :3, :]).all(-1).all(-1)
        prior[1:-1][~known] = STRUCTURE_MASK
        coords = torch.full((len(protseq) + 2, 3, 3), float("inf"))
        coords[1:-1] = coordinates[:, :3, :]
    output_dir.mkdir(parents=True, exist_ok=True)
    start_t = time()
    seq_tokens = tokenize_sequence(protseq)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)))     # unseeded like the reference: a fresh stream per call
    tokens, dt = gibbs_sample_structure_tokens(esm3_model, seq_tokens, num_samples, num_steps, temperature, top_p,
                                               prior=prior, seed=seed, structure_coords=coords)
    print(f"Sampling token time: {dt:.2f}s")
    tokens = tokens.cpu()
    if decoder is not None:
        from .decoder import decode_to_pdb
        decode_to_pdb(decoder, tokens, protseq.replace("_", "X"), save_to)
    else:
        tok_path = output_dir / f"{sample_basename}.structure_tokens.pt"
        torch.save({"sequence": protseq, "sequence_tokens": seq_tokens, "structure_tokens": tokens}, tok_path)
        print(f"no --decoder_ckpt: structure decode skipped, tokens saved to {tok_path}")
    print(f"Total time: {time() - start_t:.2f}s")
    return tokens
