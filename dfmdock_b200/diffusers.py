"""VE-SDE schedules and the one-step reverse increment (host side, fp64 scalars like the reference's numpy).

Mirrors the inference-time surface of the reference diffusers:
  SO3Diffuser.sigma / diffusion_coef / torch_reverse   src/utils/so3_diffuser.py:210-227, 344-369
  R3Diffuser.sigma / diffusion_coef / torch_reverse    src/utils/r3_diffuser.py:20-24, 40-55
The IGSO(3) tables the reference builds in __init__ (so3_diffuser.py:155-198) are training-only and are not built.
"""
import numpy as np
import torch


def _reverse_increment(g_t, score_t, dt, noise_scale, ode):
    """One Euler-Maruyama increment of the reverse VE-SDE  dx = g^2 score dt + g sqrt(dt) (noise_scale z),  z ~ N(0, I_3),
    or of its probability-flow ODE  dx = g^2 score dt / 2  (no random draw).  g_t is the fp64 host scalar g(t)."""
    g = float(g_t)
    if ode:
        return ((0.5 * (g * g)) * score_t * dt).float()
    z = noise_scale * torch.randn(1, 3, device=score_t.device)
    return ((g * g) * score_t * dt + (g * torch.sqrt(dt)) * z).float()


class _ReverseMixin:
    """torch_reverse(score_t, dt, t, noise_scale, ode) with the reference's argument order and RNG consumption
    (src/utils/so3_diffuser.py:344-369, src/utils/r3_diffuser.py:40-55): t must be a Python / numpy scalar."""

    def torch_reverse(self, score_t, dt, t, noise_scale=1.0, ode=False):
        if not np.isscalar(t):
            raise ValueError("t must be a scalar, got %r" % (t,))
        return _reverse_increment(self.diffusion_coef(t), score_t, torch.as_tensor(dt), noise_scale, ode)


class SO3Diffuser(_ReverseMixin):
    def __init__(self, conf):
        self.schedule = conf.get("schedule", "logarithmic")
        self.min_sigma = conf["min_sigma"]
        self.max_sigma = conf["max_sigma"]
        if self.schedule != "logarithmic":
            raise ValueError(f"Unrecognize schedule {self.schedule}")

    def sigma(self, t):
        if np.any(t < 0) or np.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return np.log(t * np.exp(self.max_sigma) + (1 - t) * np.exp(self.min_sigma))

    def diffusion_coef(self, t):
        s = self.sigma(t)
        return np.sqrt(2 * (np.exp(self.max_sigma) - np.exp(self.min_sigma)) * s / np.exp(s))


class R3Diffuser(_ReverseMixin):
    def __init__(self, conf):
        self.min_sigma = conf["min_sigma"]
        self.max_sigma = conf["max_sigma"]

    def sigma(self, t):
        return self.min_sigma * (self.max_sigma / self.min_sigma) ** t

    def diffusion_coef(self, t):
        return self.sigma(t) * np.sqrt(2 * (np.log(self.max_sigma) - np.log(self.min_sigma)))
