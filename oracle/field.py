"""Grid, boundary descriptor and ghost-cell (`shift`) semantics.  ORACLE (tests only).

Restates, in plain NumPy:
  * jax_ib/base/grids.py:563-740      Grid (step, cell_center, cell_faces, axes, mesh)
  * jax_ib/base/boundaries.py:95-301  ConstantBoundaryConditions.shift/_pad/_trim
  * jax_ib/base/boundaries.py:350-419 trim_boundary / pad_and_impose_bc / impose_bc

Arrays are C-order [Nx, Ny]; a Var's `offset` is in units of the grid step
(u: (1, .5), v: (.5, 1), p: (.5, .5)); coordinate of index i is
lower + (i + offset) * step (grids.py:665).
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

PERIODIC = "periodic"
DIRICHLET = "dirichlet"
NEUMANN = "neumann"


class Grid:
    """grids.py:563-637: shape + domain; step = (upper - lower) / size (Python floats)."""

    def __init__(self, shape: Sequence[int], domain: Sequence[Tuple[float, float]]):
        self.shape = tuple(int(s) for s in shape)
        self.domain = tuple((float(a), float(b)) for a, b in domain)
        self.step = tuple((b - a) / n for (a, b), n in zip(self.domain, self.shape))

    ndim = 2
    cell_center = (0.5, 0.5)
    cell_faces = ((1.0, 0.5), (0.5, 1.0))  # grids.py:632-637

    def axes(self, offset, dtype=np.float32):
        """grids.py:649-667: lower + (arange(n) + offset) * step, evaluated in `dtype`
        (JAX weak typing rounds the Python-float step/lower to float32 first)."""
        t = np.dtype(dtype).type
        return tuple(
            t(lo) + (np.arange(n, dtype=dtype) + t(o)) * t(h)
            for (lo, _), o, n, h in zip(self.domain, offset, self.shape, self.step)
        )

    def mesh(self, offset, dtype=np.float32):
        """grids.py:705-718 (indexing='ij')."""
        return tuple(np.meshgrid(*self.axes(offset, dtype), indexing="ij"))


@dataclasses.dataclass
class BC:
    """boundaries.py:40-88.  types[axis] = (lower, upper); values[axis] = (lower, upper).
    `time_stamp` is the simulation clock that rides in the velocity BC
    (boundaries.py:81); `boundary_fn` is the list of four wall functions of time."""

    types: Tuple[Tuple[str, str], ...]
    values: Tuple[Tuple[float, float], ...] = ((0.0, 0.0), (0.0, 0.0))
    time_stamp: float = 0.0
    boundary_fn: Optional[Sequence[Callable]] = None

    def with_time(self, t, values=None) -> "BC":
        return BC(self.types, self.values if values is None else values, t, self.boundary_fn)


def periodic_bc(time_stamp=0.0, boundary_fn=None) -> BC:
    """boundaries.py:852-879 new_periodic_boundary_conditions."""
    return BC(((PERIODIC, PERIODIC), (PERIODIC, PERIODIC)), ((0.0, 0.0), (0.0, 0.0)), time_stamp, boundary_fn)


def channel_bc(values=((0.0, 0.0), (0.0, 0.0)), time_stamp=0.0, boundary_fn=None) -> BC:
    """boundaries.py:636-663 Moving_wall_boundary_conditions: periodic x, Dirichlet y."""
    return BC(((PERIODIC, PERIODIC), (DIRICHLET, DIRICHLET)), values, time_stamp, boundary_fn)


def far_field_bc(values=((0.0, 0.0), (0.0, 0.0)), time_stamp=0.0, boundary_fn=None) -> BC:
    """boundaries.py:666-693 Far_field_boundary_conditions: Dirichlet on all walls."""
    return BC(((DIRICHLET, DIRICHLET), (DIRICHLET, DIRICHLET)), values, time_stamp, boundary_fn)


def pressure_bc_from_velocity(v) -> BC:
    """boundaries.py:797-809: periodic axis -> periodic, anything else -> homogeneous Neumann."""
    types = []
    for axis in range(2):
        per = {u.bc.types[axis][0] == PERIODIC for u in v}
        if len(per) != 1:
            raise ValueError("inconsistent periodicity between velocity components")
        types.append((PERIODIC, PERIODIC) if per.pop() else (NEUMANN, NEUMANN))
    return BC(tuple(types), ((0.0, 0.0), (0.0, 0.0)), 2.0, v[0].bc.boundary_fn)


def all_periodic(*vs) -> bool:
    """boundaries.py:761-767."""
    return all(u.bc.types[a][0] == PERIODIC for u in vs for a in range(2))


def _close(a, b):
    return bool(np.isclose(a, b))


def pad(data: np.ndarray, offset, grid: Grid, bc: BC, width: int, axis: int):
    """boundaries.py:116-273.  width<0 pads the lower side, width>0 the upper side.
    Returns (padded data, new offset)."""
    lower = width < 0
    n = abs(width)
    side = 0 if lower else 1
    kind = bc.types[axis][side]
    pw = [(0, 0), (0, 0)]
    pw[axis] = (n, 0) if lower else (0, n)
    new_offset = list(offset)
    new_offset[axis] -= pw[axis][0]
    if kind not in (PERIODIC, DIRICHLET) and n > 1:
        raise ValueError(f"Padding past 1 ghost cell is not defined in {kind} case.")
    two = data.dtype.type(2)

    def const(vals):
        # jnp.pad(..., mode='constant', constant_values=((lo0,hi0),(lo1,hi1)))
        cv = [(0, 0), (0, 0)]
        cv[axis] = vals
        return np.pad(data, pw, mode="constant", constant_values=cv)

    if n == 0:
        out = data
    elif kind == PERIODIC:
        out = np.pad(data, pw, mode="wrap")  # :168-176
    elif kind == DIRICHLET:
        vals = bc.values[axis]
        frac = offset[axis] % 1
        missing = grid.shape[axis] - data.shape[axis]  # >0 for a boundary-trimmed array
        if _close(frac, 0.5):  # cell centre :179-188
            out = two * const(vals) - np.pad(data, pw, mode="symmetric")
        elif _close(frac, 0.0):  # cell edge :189-246
            on_wall = (_close(offset[axis], 0) and width > 0) or (_close(offset[axis], 1) and width < 0) or missing == 1
            half = tuple(x / 2 for x in vals)
            if on_wall:
                if n == 1:  # :205-211  the first ghost point IS the wall
                    out = const(vals)
                else:  # :212-231  wall value, then (bc - reflect) beyond it
                    one = [(0, 0), (0, 0)]
                    one[axis] = (1, 0) if lower else (0, 1)
                    rest = [(0, 0), (0, 0)]
                    rest[axis] = (n - 1, 0) if lower else (0, n - 1)
                    expanded = np.pad(data, one, mode="constant", constant_values=0)
                    out = two * const(half) - np.pad(expanded, rest, mode="reflect")
            else:  # :232-246  ghost = bc - reflect(interior)
                out = two * const(half) - np.pad(data, pw, mode="reflect")
        else:
            raise ValueError(f"expected offset to be an edge or cell center, got {offset[axis]}")
    elif kind == NEUMANN:  # :247-269  ghost = edge + h * (0 - bc)
        h = data.dtype.type(grid.step[axis])
        out = np.pad(data, pw, mode="edge") + h * (np.pad(data, pw, mode="constant") - const(bc.values[axis]))
    else:
        raise ValueError("invalid boundary type")
    return out, tuple(new_offset)


def trim(data: np.ndarray, offset, width: int, axis: int):
    """boundaries.py:275-301."""
    lo, hi = (-width, 0) if width < 0 else (0, width)
    sl = [slice(None), slice(None)]
    sl[axis] = slice(lo, data.shape[axis] - hi)
    new_offset = list(offset)
    new_offset[axis] += lo
    return data[tuple(sl)], tuple(new_offset)


@dataclasses.dataclass
class Var:
    """GridVariable (grids.py:331-458): data + offset + grid + bc."""

    data: np.ndarray
    offset: Tuple[float, float]
    grid: Grid
    bc: BC

    def shift(self, k: int, axis: int) -> "Var":
        """boundaries.py:95-114: shift = trim(pad(u, k), -k); result offset += k."""
        padded, off = pad(self.data, self.offset, self.grid, self.bc, k, axis)
        data, off = trim(padded, off, -k, axis)
        return Var(data, off, self.grid, self.bc)

    def with_data(self, data, offset=None, bc=None) -> "Var":
        return Var(data, self.offset if offset is None else offset, self.grid, self.bc if bc is None else bc)

    def impose_bc(self) -> "Var":
        """boundaries.py:350-419: re-stamp wall values on edge-aligned Dirichlet faces."""
        data, off = self.data, self.offset
        for axis in range(2):  # trim_boundary :350-373
            if _close(off[axis], 0.0) and self.bc.types[axis][0] == DIRICHLET:
                data, off = trim(data, off, -1, axis)
            elif _close(off[axis], 1.0) and self.bc.types[axis][1] == DIRICHLET:
                data, off = trim(data, off, 1, axis)
        for axis in range(2):  # pad_and_impose_bc :375-403
            if self.bc.types[axis][0] == DIRICHLET and _close(off[axis], 1.0) and data.shape[axis] < self.grid.shape[axis]:
                if _close(self.offset[axis], 1.0):
                    data, off = pad(data, off, self.grid, self.bc, 1, axis)
                elif _close(self.offset[axis], 0.0):
                    data, off = pad(data, off, self.grid, self.bc, -1, axis)
        return Var(data, off, self.grid, self.bc)
