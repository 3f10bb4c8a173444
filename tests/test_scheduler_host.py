"""Product-side host scheduler (dreamlab_b200.scheduler) against the oracle: integer schedule
bit-exact, fp32 coefficients bit-identical to the oracle's per-step scalars."""
import pytest
import torch

from oracle.scheduler import OracleLCMScheduler, guidance_scale_embedding as oracle_gse
import dreamlab_b200.scheduler as S


@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 8, 16, 30, 50])
def test_timesteps_and_coefficients(n):
    o = OracleLCMScheduler()
    ts = o.set_timesteps(n)
    s = S.LCMSchedule(n)
    assert s.timesteps == ts.tolist()
    for i, t in enumerate(ts.tolist()):
        prev = ts[i + 1].item() if i + 1 < n else t
        a_t, a_p = o.alphas_cumprod[t], o.alphas_cumprod[prev]
        c_skip, c_out = o.boundary_scalings(t)
        want = (a_t.sqrt().item(), (1 - a_t).sqrt().item(), float(c_skip), float(c_out),
                a_p.sqrt().item(), (1 - a_p).sqrt().item())
        assert s.coeffs(i) == want
        assert s.has_noise(i) == (i != n - 1)


def test_bad_step_counts():
    for n in (0, -1, 51):
        with pytest.raises(ValueError):
            S.LCMSchedule(n)


def test_guidance_embedding_matches_oracle():
    w = torch.tensor([0.0, 0.5, 6.5])
    assert torch.equal(S.guidance_scale_embedding(w, 256), oracle_gse(w, 256))
    assert S.guidance_scale_embedding(w, 33).shape == (3, 33)
