"""Predefined noise schedules (src/noise.py:7-56, 92-128): a float64 numpy table of gamma = -log(alpha^2/sigma^2),
stored fp32 under the same state_dict key (`gamma`) as the reference."""
import numpy as np
import torch


def _clip_ratio(alphas2, clip_value=0.001):
    """alpha_t^2/alpha_{t-1}^2 clipped to [clip, 1], re-accumulated (src/noise.py:7-19)."""
    ext = np.concatenate([np.ones(1), alphas2], axis=0)
    return np.cumprod(np.clip(ext[1:] / ext[:-1], a_min=clip_value, a_max=1.0), axis=0)


def polynomial_schedule(timesteps: int, s=1e-4, power=3.0):
    """src/noise.py:22-36."""
    steps = timesteps + 1
    grid = np.linspace(0, steps, steps)
    alphas2 = _clip_ratio((1 - np.power(grid / steps, power)) ** 2)
    return (1 - 2 * s) * alphas2 + s


def cosine_beta_schedule(timesteps, s=0.008, raise_to_power: float = 1):
    """src/noise.py:39-56."""
    steps = timesteps + 2
    grid = np.linspace(0, steps, steps)
    ac = np.cos(((grid / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)
    ac = np.cumprod(1.0 - betas, axis=0)
    return np.power(ac, raise_to_power) if raise_to_power != 1 else ac


class PredefinedNoiseSchedule(torch.nn.Module):
    def __init__(self, noise_schedule, timesteps, precision):
        super().__init__()
        self.timesteps = timesteps
        if noise_schedule == "cosine":
            alphas2 = cosine_beta_schedule(timesteps)
        elif "polynomial" in noise_schedule:
            parts = noise_schedule.split("_")
            assert len(parts) == 2
            alphas2 = polynomial_schedule(timesteps, s=precision, power=float(parts[1]))
        else:
            raise ValueError(noise_schedule)
        gamma = -(np.log(alphas2) - np.log(1 - alphas2))
        self.gamma = torch.nn.Parameter(torch.from_numpy(gamma).float(), requires_grad=False)

    def forward(self, t):
        """gamma[round(t * timesteps)] -- `timesteps` is the table's own length even if EDM.T was overridden
        (src/noise.py:126-128; generate.py:103-104)."""
        return self.gamma[torch.round(t * self.timesteps).long()]
