"""Advection plug-ins with the reference's names and signatures (jax_ib/base/advection.py:120-124,
206-280, 325-354), computed by the CUDA stencil kernel (csrc/stencil*.cu) through the C ABI.

    convect = lambda v: tuple(advection.advect_upwind(u, v, dt) for u in v)     # notebook cell 3

Each call returns -div(c u) for ONE transported component `c` (a velocity component of `v`; the
kernel evaluates both components in one pass and the requested one is returned).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _probe, engine, grids, ops

GridArray = grids.GridArray
GridVariable = grids.GridVariable


def _advect(scheme: str, c, v, dt):
    if _probe.is_probe(c) or _probe.is_probe(v):
        return _probe.Tag("convect", scheme)
    offs = [tuple(u.offset) for u in v]
    if tuple(c.offset) not in offs or c.data.data_ptr() != v[offs.index(tuple(c.offset))].data.data_ptr():
        raise NotImplementedError("the CUDA advection kernels transport the velocity components themselves "
                                  "(c must be one of v); passive scalars are not on the jax_ib hot path")
    grid = grids.consistent_grid(c, *v)
    spec = engine.grid_spec_from(grid, dt if dt is not None else 0.0, None, 1.0, engine.bc_kind_of(v), scheme, v)
    du, dv = ops.explicit_terms(spec, *(engine.to_device(u.data) for u in v))
    out = (du, dv)[offs.index(tuple(c.offset))]
    return GridArray(out, c.offset, grid)


def advect_upwind(c: GridVariable, v, dt: Optional[float] = None) -> GridArray:
    """advection.py:120-124."""
    return _advect("upwind", c, v, dt)


def advect_van_leer(c: GridVariable, v, dt: float) -> GridArray:
    """advection.py:206-280."""
    return _advect("van_leer", c, v, dt)


def advect_van_leer_using_limiters(c: GridVariable, v, dt: float) -> GridArray:
    """advection.py:325-333 (the default `convect`, equations.py:55-58)."""
    return _advect("van_leer_limiters", c, v, dt)


def stable_time_step(max_velocity: float, max_courant_number: float, grid: grids.Grid) -> float:
    """advection.py:336-354."""
    return max_courant_number * min(grid.step) / max_velocity
