"""Boundary-condition objects mirrored from the reference's jax_ib/base/boundaries.py: the
time-stamped ConstantBoundaryConditions (:40-422) that carries the simulation clock, the BC factories
(:546-693, :852-879), update_BC / Reserve_BC (:492-542) and the pressure-BC inference (:797-809).

`shift` is provided for user-level post-processing with torch ops (periodic wrap, Dirichlet and
Neumann ghost cells for one or two ghost layers); the CUDA kernels implement the same ghost rules
internally and never call this.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import grids
from . import particle_class

GridArray = grids.GridArray
GridVariable = grids.GridVariable


class BCType:
    PERIODIC = "periodic"
    DIRICHLET = "dirichlet"
    NEUMANN = "neumann"


def _take(data: torch.Tensor, axis: int, idx) -> torch.Tensor:
    return data.index_select(axis, torch.as_tensor(idx, device=data.device, dtype=torch.long))


class ConstantBoundaryConditions:
    """boundaries.py:40-422.  Constructor argument order follows the reference:
    (time_stamp, values, types, boundary_fn)."""

    def __init__(self, time_stamp, values, types, boundary_fn):
        self.types = tuple(tuple(t) for t in types)
        self.bc_values = tuple(tuple(v) for v in values)
        self.boundary_fn = boundary_fn
        self.time_stamp = time_stamp if time_stamp is not None else []

    # -- ghost cells --------------------------------------------------------------------------
    def _ghost(self, u: GridArray, k: int, axis: int, lower: bool) -> torch.Tensor:
        """Value of the k-th ghost layer (k = 1 nearest) on one side (boundaries.py:116-273)."""
        d = u.data
        n = d.shape[axis]
        kind = self.types[axis][0 if lower else 1]
        if kind == BCType.PERIODIC:
            return _take(d, axis, [n - k] if lower else [k - 1])
        bc = self.bc_values[axis][0 if lower else 1]
        if kind == BCType.DIRICHLET:
            frac = u.offset[axis] % 1
            if np.isclose(frac, 0.5):
                return 2 * bc - _take(d, axis, [k - 1] if lower else [n - k])
            if np.isclose(frac, 0.0):
                on_wall = (np.isclose(u.offset[axis], 1) and lower) or (np.isclose(u.offset[axis], 0) and not lower)
                if on_wall:
                    if k == 1:
                        return torch.full_like(_take(d, axis, [0]), bc)
                    return bc - _take(d, axis, [k - 2] if lower else [n - k + 1])
                return bc - _take(d, axis, [k] if lower else [n - 1 - k])
            raise ValueError(f"expected offset to be an edge or cell center, got {u.offset[axis]}")
        if kind == BCType.NEUMANN:
            if k > 1:
                raise ValueError("Padding past 1 ghost cell is not defined in neumann case.")
            h = u.grid.step[axis]
            return _take(d, axis, [0] if lower else [n - 1]) - h * bc
        raise ValueError("invalid boundary type")

    def shift(self, u: GridArray, offset: int, axis: int) -> GridArray:
        """boundaries.py:95-114."""
        if offset == 0:
            return u
        d = u.data
        n = d.shape[axis]
        k = abs(offset)
        new_offset = tuple(o + offset if a == axis else o for a, o in enumerate(u.offset))
        if offset > 0:
            ghosts = [self._ghost(u, g, axis, lower=False) for g in range(1, k + 1)]
            out = torch.cat([d.narrow(axis, k, n - k)] + ghosts, dim=axis)
        else:
            ghosts = [self._ghost(u, g, axis, lower=True) for g in range(k, 0, -1)]
            out = torch.cat(ghosts + [d.narrow(axis, 0, n - k)], dim=axis)
        return GridArray(out, new_offset, u.grid)

    def trim_boundary(self, u: GridArray) -> GridArray:
        """boundaries.py:350-373."""
        d, off = u.data, list(u.offset)
        for axis in range(u.grid.ndim):
            if np.isclose(off[axis], 0.0) and self.types[axis][0] == BCType.DIRICHLET:
                d = d.narrow(axis, 1, d.shape[axis] - 1)
                off[axis] += 1
            elif np.isclose(off[axis], 1.0) and self.types[axis][1] == BCType.DIRICHLET:
                d = d.narrow(axis, 0, d.shape[axis] - 1)
        return GridArray(d, tuple(off), u.grid)

    def impose_bc(self, u: GridArray) -> GridVariable:
        """boundaries.py:405-419: re-stamp wall values on edge-aligned Dirichlet faces."""
        d = u.data.clone()
        for axis in range(u.grid.ndim):
            n = d.shape[axis]
            if np.isclose(u.offset[axis], 1.0) and self.types[axis][1] == BCType.DIRICHLET:
                d.narrow(axis, n - 1, 1).fill_(self.bc_values[axis][1])
            elif np.isclose(u.offset[axis], 0.0) and self.types[axis][0] == BCType.DIRICHLET:
                d.narrow(axis, 0, 1).fill_(self.bc_values[axis][0])
        return GridVariable(GridArray(d, u.offset, u.grid), self)

    def values(self, axis: int, grid: grids.Grid):
        if None in self.bc_values[axis]:
            return (None, None)
        shape = grid.shape[:axis] + grid.shape[axis + 1:]
        return tuple(torch.full(shape, float(self.bc_values[axis][-i])) for i in [0, 1])


class HomogeneousBoundaryConditions(ConstantBoundaryConditions):
    """boundaries.py:424-447."""

    def __init__(self, types):
        super().__init__(0.0, ((0.0, 0.0),) * len(types), types, lambda x: x)


def periodic_boundary_conditions(ndim: int):
    return HomogeneousBoundaryConditions(((BCType.PERIODIC, BCType.PERIODIC),) * ndim)


def channel_flow_boundary_conditions(ndim: int, bc_vals=None):
    """boundaries.py:598-631."""
    types = ((BCType.PERIODIC, BCType.PERIODIC), (BCType.DIRICHLET, BCType.DIRICHLET)) + \
        ((BCType.PERIODIC, BCType.PERIODIC),) * (ndim - 2)
    if not bc_vals:
        return HomogeneousBoundaryConditions(types)
    return ConstantBoundaryConditions(0.0, bc_vals, types, lambda x: x)


def new_periodic_boundary_conditions(ndim, bc_vals, time_stamp, bc_fn):
    """boundaries.py:852-879."""
    types = ((BCType.PERIODIC, BCType.PERIODIC),) * ndim
    return ConstantBoundaryConditions(values=bc_vals, time_stamp=time_stamp, types=types, boundary_fn=bc_fn)


def Moving_wall_boundary_conditions(ndim, bc_vals, time_stamp, bc_fn):
    """boundaries.py:636-663: periodic x, Dirichlet y."""
    types = ((BCType.PERIODIC, BCType.PERIODIC), (BCType.DIRICHLET, BCType.DIRICHLET)) + \
        ((BCType.PERIODIC, BCType.PERIODIC),) * (ndim - 2)
    return ConstantBoundaryConditions(values=bc_vals, time_stamp=time_stamp, types=types, boundary_fn=bc_fn)


def Far_field_boundary_conditions(ndim, bc_vals, time_stamp, bc_fn):
    """boundaries.py:666-693: Dirichlet on every wall (no CUDA path for this family yet)."""
    types = ((BCType.DIRICHLET, BCType.DIRICHLET),) * ndim
    return ConstantBoundaryConditions(values=bc_vals, time_stamp=time_stamp, types=types, boundary_fn=bc_fn)


def is_periodic_boundary_conditions(c: GridVariable, axis: int) -> bool:
    return c.bc.types[axis][0] == BCType.PERIODIC


def has_all_periodic_boundary_conditions(*arrays: GridVariable) -> bool:
    """boundaries.py:761-767."""
    return all(is_periodic_boundary_conditions(a, ax) for a in arrays for ax in range(a.grid.ndim))


def consistent_boundary_conditions(*arrays: GridVariable):
    """boundaries.py:770-794."""
    out = []
    for axis in range(arrays[0].grid.ndim):
        bcs = {is_periodic_boundary_conditions(a, axis) for a in arrays}
        if len(bcs) != 1:
            raise grids.InconsistentBoundaryConditionsError(f"arrays do not have consistent bc: {arrays}")
        out.append("periodic" if bcs.pop() else "nonperiodic")
    return tuple(out)


def get_pressure_bc_from_velocity(v):
    """boundaries.py:797-809."""
    types = [(BCType.PERIODIC, BCType.PERIODIC) if t == "periodic" else (BCType.NEUMANN, BCType.NEUMANN)
             for t in consistent_boundary_conditions(*v)]
    return ConstantBoundaryConditions(values=((0.0, 0.0), (0.0, 0.0)), time_stamp=2.0, types=types,
                                      boundary_fn=v[0].bc.boundary_fn)


def _advance(all_variable, step_time, reserve: bool):
    v = all_variable.velocity
    fx, fy = v[0].bc.boundary_fn, v[1].bc.boundary_fn
    ts = np.float32(v[0].bc.time_stamp) + np.float32(step_time)
    t2 = 0.0 if reserve else ts
    vx_bc = ((fx[0](ts), fx[1](t2)), (fx[2](ts), fx[3](t2)))
    vy_bc = ((fy[0](ts), fy[1](t2)), (fy[2](ts), fy[3](t2)))
    bcs = (ConstantBoundaryConditions(values=vx_bc, time_stamp=ts, types=v[0].bc.types, boundary_fn=fx),
           ConstantBoundaryConditions(values=vy_bc, time_stamp=ts, types=v[1].bc.types, boundary_fn=fy))
    v_new = tuple(GridVariable(u.array, bc) for u, bc in zip(v, bcs))
    return dataclasses.replace(all_variable, velocity=v_new)


def update_BC(all_variable, step_time: float):
    """boundaries.py:519-542: time_stamp += dt (float32) and re-evaluate the four wall functions."""
    return _advance(all_variable, step_time, reserve=False)


def Reserve_BC(all_variable, step_time: float):
    """boundaries.py:492-515."""
    return _advance(all_variable, step_time, reserve=True)
