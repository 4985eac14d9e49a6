"""jax_ib/MD/CFD_MD_coupling.py:4-27: bilinear E->L sampling of (u, v) at tracer positions
(map_coordinates order=1, zero outside the grid), by the CUDA kernel csrc/tracer.cu."""
from __future__ import annotations

import torch

from .. import engine, ops


def interpolate_pbc(field, list_p):
    """CFD_MD_coupling.py:7-19: index = x/dx - offset; despite the name there is no periodic wrap."""
    grid = field.grid
    spec = engine.grid_spec_from(grid, 0.0, None, 1.0, "periodic", "none")
    return ops.bilinear_sample(spec, engine.to_device(field.data), field.offset, engine.to_device(list_p))


def custom_force_fn_pbc(all_variables):
    """CFD_MD_coupling.py:22-27."""
    v = all_variables.velocity
    R = all_variables.MD_var.position
    return torch.stack([interpolate_pbc(v[0], R), interpolate_pbc(v[1], R)], dim=-1) * 1.0
