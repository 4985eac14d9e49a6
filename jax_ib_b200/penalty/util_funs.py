"""jax_ib/penalty/util_funs.py:6-123: permeability mask of parametrised bodies and the penalty forcing
`grad_p - kappa * u`.  The mask is evaluated by the CUDA kernel csrc/tracer.cu (perm_kernel) with the
tanh smoothing K/2 (1 + tanh(G / width)); the forcing fuses into the predictor kernel."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine, grids, ops


def tanh_smoothing(K: float, width: float):
    """A `smoothening_fn(G, Know)` for calc_perm: K/2 (1 + tanh(G / width)) (util_funs.py:91)."""

    def fn(G, Know=None):
        return K / 2.0 * (1.0 + torch.tanh(G / width))

    fn._jaxib_tanh = (float(K), float(width))
    return fn


def perm_vmap_multiple_particles(grid, particles, smoothening_fn, Know=None):
    """util_funs.py:95-123: sum over bodies of smoothening_fn(R(theta)^2 - r^2)."""
    if not hasattr(smoothening_fn, "_jaxib_tanh"):
        raise NotImplementedError("the CUDA mask kernel implements util_funs.tanh_smoothing(K, width)")
    K, width = smoothening_fn._jaxib_tanh
    plist = particles if isinstance(particles, tuple) else (particles,)
    spec = engine.grid_spec_from(grid, 0.0, None, 1.0, "periodic", "none")
    perm = None
    for p in plist:
        theta = p.generate_grid()
        Rt = np.stack([np.asarray(p.shape(g, theta), dtype=np.float32) for g in p.geometry_param])
        centers = np.asarray(p.particle_center, dtype=np.float32).reshape(-1, 2)
        out = ops.penalty_perm(spec, engine.to_device(centers), engine.to_device(Rt), K, width)
        perm = out if perm is None else perm + out
    return grids.GridArray(perm, grid.cell_center, grid)


def arbitrary_obstacle(pressure_gradient, permeability):
    """util_funs.py:6-16: forcing(v) = (px - kappa u, py - kappa v)."""
    perm = permeability.data if hasattr(permeability, "data") else permeability

    def forcing(v):
        return tuple(grids.GridArray(g * torch.ones_like(u.data) - engine.to_device(perm) * u.data, u.offset, u.grid)
                     for g, u in zip(pressure_gradient, v))

    forcing._jaxib_penalty = ((float(pressure_gradient[0]), float(pressure_gradient[1])), perm)
    return forcing
