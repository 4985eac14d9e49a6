"""jax_ib/MD/simulate.py:41-99 brownian_cfd: over-damped Brownian integrator driven by the CFD velocity.

    init_fn, apply_fn = brownian_cfd(energy_or_force, shift, dt, kT, CFD_force, gamma)
    state = init_fn(key, R, mass=1.)          # BrownianState(position, mass, rng)
    state = apply_fn(all_variables, **kwargs) # reads all_variables.MD_var

kT = 0 is deterministic and matches the reference's arithmetic.  For kT > 0 the deviates come from a
torch Generator seeded with `key` (JAX's threefry stream is not reproduced: statistical parity only).
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable

import numpy as np
import torch

from .. import engine


@dataclasses.dataclass
class BrownianState:
    position: Any
    mass: Any
    rng: Any


def brownian_cfd(energy_or_force: Callable, shift: Callable, dt: float, kT: float, CFD_force: Callable, gamma: float = 0.1):
    def init_fn(key, R, mass=1.0):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(np.asarray(key).ravel()[-1]) if not isinstance(key, int) else key)
        return BrownianState(engine.to_device(R), float(mass), gen)

    def apply_fn(all_variables, **kwargs):
        state = all_variables.MD_var
        _kT = kwargs.get("kT", kT)
        R = state.position
        F = energy_or_force(R, **kwargs) + CFD_force(all_variables)
        nu = 1.0 / (state.mass * gamma)
        dR = F * np.float32(dt) * np.float32(nu)
        if _kT:
            xi = torch.randn(R.shape, generator=state.rng, dtype=torch.float32).to(R.device)
            dR = dR + float(np.sqrt(np.float32(2.0) * np.float32(_kT) * np.float32(dt) * np.float32(nu))) * xi
        return BrownianState(shift(R, dR, **kwargs), state.mass, state.rng)

    return init_fn, apply_fn
