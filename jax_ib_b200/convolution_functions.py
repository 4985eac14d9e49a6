"""Discrete delta and E->L interpolation with the reference's names
(jax_ib/base/convolution_functions.py:5-37), evaluated by the CUDA kernel csrc/ib.cu."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _probe, engine, grids, ops


def delta_approx_logistjax(x, x0, w):
    """convolution_functions.py:5-7: Gaussian of width w (always the grid step in the reference)."""
    if _probe.is_probe(x) or _probe.is_probe(x0):
        return _probe.Tag("delta", "gaussian")
    x = torch.as_tensor(x, dtype=torch.float32)
    x0 = torch.as_tensor(x0, dtype=torch.float32)
    return 1 / (w * math.sqrt(2 * math.pi)) * torch.exp(-0.5 * ((x - x0) / w) ** 2)


def new_surf_fn(field, xp, yp, discrete_fn, window: int = 6):
    """convolution_functions.py:11-37: U_k = sum_ij field * delta(xp-X) * delta(yp-Y) * dx * dy.
    Only the Gaussian delta above is supported (the kernel truncates it at `window` cells, 6e-9)."""
    tag = _probe.try_probe(discrete_fn, _probe.Probe(), 0.0, 1.0)
    if not (isinstance(tag, _probe.Tag) and tag.value == "gaussian"):
        raise NotImplementedError("new_surf_fn: the CUDA path implements delta_approx_logistjax only")
    if _probe.is_probe(field) or _probe.is_probe(xp):
        return _probe.Tag("surface", "gaussian", window)
    grid = field.grid
    spec = engine.grid_spec_from(grid, 0.0, None, 1.0, "periodic", "none", ib_window=window)
    xp = engine.to_device(torch.as_tensor(np.asarray(xp, dtype=np.float32)) if not isinstance(xp, torch.Tensor) else xp)
    yp = engine.to_device(torch.as_tensor(np.asarray(yp, dtype=np.float32)) if not isinstance(yp, torch.Tensor) else yp)
    return ops.ib_interp(spec, engine.to_device(field.data), field.offset, xp, yp)
