"""CPU oracle for the ESMDiff ddpm sampling path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or as the timed CPU baseline) -- never
as the thing shipped.  The product path (``esmdiff_b200``) fails loudly when the CUDA
extension is missing; it never routes through this package.

Contents
--------
* ``esm3_ref``   fp32 PyTorch restatement of the ESM3-open trunk as used by
                 ``CustomizedESM3.forward`` (reference slm/models/net.py:322-483) and of the
                 ``esm==3.0.4`` layers it calls (third-party, absent from /root/reference:
                 **parity unpinned** for that half -- see the module docstring).
* ``mdlm_ref``   restatement of the MDLM sampler (reference slm/models/model.py:24-28,
                 464-492, 527-607) and LogLinearNoise (slm/utils/noise_utils.py:188-213).
                 Pinned bit-for-bit against the reference's own code run in the build
                 container (``ref_loader`` + ``make_golden`` -> ``tests/golden/``).
* ``geom_ref`` / ``vqvae_enc_ref`` / ``vqvae_ref`` / ``gibbs_ref``  restatements of the esm==3.0.4 pieces around the
                 path: backbone frames + geometric attention, the VQ-VAE structure encoder and decoder, the
                 iterative structure sampler (**parity unpinned**, each says so in its header).
* ``ref_loader`` imports the reference's model.py / noise_utils.py verbatim through stub
                 modules.  Works only where /root/reference exists (the build container).
* ``make_golden_model_step`` the reference's own ``model_step`` (forward half) -> ``tests/golden/model_step.npz``.
* ``make_golden`` regenerates ``tests/golden/*.npz`` from the verbatim reference.
"""
