"""Oracle LCMScheduler (test infrastructure).

Restates diffusers' `LCMScheduler` as the reference uses it:
`pipe.scheduler = LCMScheduler.from_config(...)` (`backends/cuda_worker.py:88`),
`scheduler.set_timesteps(n)` / `scheduler.step(noise_pred, t, latents)`
(`backends/rknnlcm.py:559-560`, `:596-598`).  Defaults per SURVEY.md App. A.5.
"""
from __future__ import annotations

import numpy as np
import torch


class OracleLCMScheduler:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, original_inference_steps: int = 50,
                 timestep_scaling: float = 10.0):
        self.num_train_timesteps = num_train_timesteps
        self.original_inference_steps = original_inference_steps
        self.timestep_scaling = timestep_scaling
        self.sigma_data = 0.5
        # scaled_linear schedule, fp32 exactly as diffusers builds it
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                               dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)  # set_alpha_to_one=True
        self.init_noise_sigma = 1.0
        self.timesteps = None
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int):
        k = self.num_train_timesteps // self.original_inference_steps
        origin = np.asarray(list(range(1, self.original_inference_steps + 1))) * k - 1
        origin = origin[::-1].copy()
        idx = np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)
        idx = np.floor(idx).astype(np.int64)
        self.timesteps = torch.from_numpy(origin[idx]).to(torch.int64)
        self.num_inference_steps = num_inference_steps
        self._step_index = None
        return self.timesteps

    def boundary_scalings(self, t):
        # diffusers' get_scalings_for_boundary_condition_discrete receives the 0-dim int64
        # tensor the pipeline iterates over, so this arithmetic runs in fp32 tensor math
        s = torch.as_tensor(int(t), dtype=torch.int64) * self.timestep_scaling
        c_skip = self.sigma_data ** 2 / (s ** 2 + self.sigma_data ** 2)
        c_out = s / (s ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor,
             generator=None, noise: torch.Tensor | None = None):
        """Returns (prev_sample, denoised).  `noise` (if given) replaces the
        randn draw so a test can feed the CUDA path the very same z."""
        if self._step_index is None:
            self._step_index = int((self.timesteps == int(timestep)).nonzero()[0].item())
        i = self._step_index
        n = len(self.timesteps)
        prev_t = int(self.timesteps[i + 1]) if i + 1 < n else int(timestep)
        a_t = self.alphas_cumprod[int(timestep)]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        c_skip, c_out = self.boundary_scalings(int(timestep))
        x0 = (sample - b_t.sqrt() * model_output) / a_t.sqrt()
        denoised = c_out * x0 + c_skip * sample
        if i != n - 1:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator,
                                    dtype=denoised.dtype)
            prev = a_prev.sqrt() * denoised + b_prev.sqrt() * noise
        else:
            prev = denoised
        self._step_index += 1
        return prev, denoised


def guidance_scale_embedding(w: torch.Tensor, embedding_dim: int = 256) -> torch.Tensor:
    """`get_guidance_scale_embedding`, semantics of `backends/rknnlcm.py:651-677`.
    `w` is already `guidance_scale - 1` (`rknnlcm.py:574`)."""
    w = w.to(torch.float32) * 1000.0
    half = embedding_dim // 2
    e = np.log(10000.0) / (half - 1)
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    emb = w[:, None] * f[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1))
    return emb
