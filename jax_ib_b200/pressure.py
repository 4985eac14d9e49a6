"""Pressure projection plug-ins with the reference's names (jax_ib/base/pressure.py:58-170), computed
by the hand-written FFT kernels (csrc/poisson*.cu)."""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import torch

from . import _probe, boundaries, engine, grids, ops


def _solve(v, kind: str):
    grid = grids.consistent_grid(*v)
    spec = engine.grid_spec_from(grid, 0.0, None, 1.0, kind, "none", v)
    plan = engine.get_plan(spec)
    q = plan.pressure_solve(engine.to_device(v[0].data), engine.to_device(v[1].data))
    return grids.GridArray(q, grid.cell_center, grid)


def solve_fast_diag(v, q0=None, implementation: Optional[str] = None) -> grids.GridArray:
    """pressure.py:94-108: periodic x periodic, rfft route of jax_cfd fast_diagonalization."""
    if _probe.is_probe(v):
        return _probe.Tag("pressure_solve", "periodic")
    if not boundaries.has_all_periodic_boundary_conditions(*v):
        raise ValueError("solve_fast_diag() expects periodic velocity BC")
    return _solve(v, "periodic")


def solve_fast_diag_moving_wall(v, q0=None, implementation: Optional[str] = "matmul") -> grids.GridArray:
    """pressure.py:111-130: periodic x, Neumann y (the reference's eigh + matmul route == FFT x DCT-II)."""
    if _probe.is_probe(v):
        return _probe.Tag("pressure_solve", "channel")
    return _solve(v, "channel")


def solve_fast_diag_Far_Field(v, q0=None, implementation=None):
    """pressure.py:134-154: Neumann x Neumann pressure (Dirichlet velocity on all four walls).  The reference's
    eigh + matmul route over laplacian_matrix_w_boundaries (array_utils.py:254-298) == DCT-II on both axes."""
    if _probe.is_probe(v):
        return _probe.Tag("pressure_solve", "far_field")
    return _solve(v, "far_field")


def projection_and_update_pressure(All_variables, solve: Callable = solve_fast_diag):
    """pressure.py:58-91: q = solve(v); p += q; v -= forward_difference(q) [+ impose_bc]."""
    v = All_variables.velocity
    tag = _probe.try_probe(solve, _probe.velocity_probe(), None)
    kinds = _probe.tags_of(tag, "pressure_solve") if tag is not None else None
    grid = grids.consistent_grid(*v)
    pbc = boundaries.get_pressure_bc_from_velocity(v)
    if kinds:
        spec = engine.grid_spec_from(grid, 0.0, None, 1.0, kinds[0].value, "none", v)
        plan = engine.get_plan(spec)
        u, w, p = plan.project(engine.to_device(v[0].data), engine.to_device(v[1].data),
                               engine.to_device(All_variables.pressure.data))
        new_v = tuple(grids.GridVariable(grids.GridArray(d, x.offset, grid), x.bc) for d, x in zip((u, w), v))
        new_p = grids.GridVariable(grids.GridArray(p, grid.cell_center, grid), pbc)
        return dataclasses.replace(All_variables, velocity=new_v, pressure=new_p)
    # opaque user solver: compose on the device with torch ops
    from . import finite_differences as fd

    q0 = grids.GridVariable(grids.GridArray(torch.zeros(grid.shape, device=v[0].data.device), grid.cell_center, grid), pbc)
    qsol = solve(v, q0)
    q = grids.GridVariable(qsol, pbc)
    new_p = grids.GridVariable(grids.GridArray(qsol.data + All_variables.pressure.data, qsol.offset, qsol.grid), pbc)
    q_grad = fd.forward_difference(q)
    periodic = boundaries.has_all_periodic_boundary_conditions(*v)
    new_v = []
    for u, g in zip(v, q_grad):
        nv = grids.GridVariable(u.array - g, u.bc)
        new_v.append(nv if periodic else nv.impose_bc())
    return dataclasses.replace(All_variables, velocity=tuple(new_v), pressure=new_p)


def calc_P(v, solve: Callable = solve_fast_diag):
    """pressure.py:156-170."""
    pbc = boundaries.get_pressure_bc_from_velocity(v)
    return grids.GridVariable(solve(v, None), pbc)
