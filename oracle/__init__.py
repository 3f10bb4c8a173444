"""CPU oracle for the Dream Lab hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

A plain PyTorch fp32 restatement of what `backends/cuda_worker.py:221-229`
(`self.pipe(...)`) computes through the third-party `diffusers` package:
`LCMScheduler`, `UNet2DConditionModel` (SD1.5 + LCM `time_cond_proj_dim=256`),
the `AutoencoderKL` decoder and the `StableDiffusionPipeline` loop mirrored in
the reference's own `backends/rknnlcm.py:523-647`.

PARITY UNPINNED at the diffusers boundary: `diffusers` is an *unpinned*
dependency of the reference (`requirements.txt:31`), it is absent from this
image and cannot be installed (no network), and the reference's tests hold no
golden tensors for this path.  The oracle is therefore pinned
only by the known answers that are derivable offline: the LCM timestep tables,
alphas_cumprod values, c_skip/c_out, the w=0 guidance embedding and the exact
published parameter counts (859.60 M / 49.49 M) — see tests/test_oracle_kat.py.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product path
(`dreamlab_b200`, `backends/b200_worker.py`) never does.
"""
