"""Diffusion plug-in with the reference's name and signature (jax_ib/base/diffusion.py:35-57)."""
from __future__ import annotations

from . import _probe, engine, grids, ops


def diffuse(c: grids.GridVariable, nu: float) -> grids.GridArray:
    """diffusion.py:35-37: nu * laplacian(c), for a velocity component (offset (1,.5) or (.5,1))."""
    if _probe.is_probe(c):
        return _probe.Tag("diffuse", "laplacian")
    grid = c.grid
    faces = grid.cell_faces
    if tuple(c.offset) not in faces:
        raise NotImplementedError("the CUDA Laplacian handles the staggered velocity offsets only")
    comp = faces.index(tuple(c.offset))
    kind = "periodic" if c.bc.types[1][0] == "periodic" else "channel"
    spec = engine.grid_spec_from(grid, 0.0, nu, 1.0, kind, "none", (c, c) if kind == "channel" else None)
    data = engine.to_device(c.data)
    du, dv = ops.explicit_terms(spec, data, data)
    return grids.GridArray((du, dv)[comp], c.offset, grid)


def stable_time_step(viscosity: float, grid: grids.Grid) -> float:
    """diffusion.py:40-57."""
    if viscosity == 0:
        return float("inf")
    return min(grid.step) ** 2 / (viscosity * 2 ** grid.ndim)
