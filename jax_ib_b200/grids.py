"""Data model mirrored from the reference's jax_ib/base/grids.py (Grid :563-740, GridArray :107-177,
GridVariable :331-458) with torch tensors as the array type.  Pure plumbing: the hot path never
goes through these Python-level operators; they exist so that user code written against the
reference (initial conditions, post-processing, custom plug-ins) keeps working.
"""
from __future__ import annotations

import dataclasses
import numbers
import operator
from typing import Any, Callable, Optional, Sequence, Tuple

import numpy as np
import torch


class InconsistentOffsetError(Exception):
    """grids.py:520-525."""


class InconsistentGridError(Exception):
    """grids.py:527-532."""


class InconsistentBoundaryConditionsError(Exception):
    """grids.py:534-560."""


def _to_tensor(x, like: Optional[torch.Tensor] = None):
    if isinstance(x, torch.Tensor):
        return x
    dev = like.device if like is not None else "cpu"
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=dev)


@dataclasses.dataclass(init=False, frozen=True)
class Grid:
    """grids.py:563-740."""

    shape: Tuple[int, ...]
    step: Tuple[float, ...]
    domain: Tuple[Tuple[float, float], ...]

    def __init__(self, shape, step=None, domain=None):
        shape = tuple(operator.index(s) for s in shape)
        object.__setattr__(self, "shape", shape)
        if step is not None and domain is not None:
            raise TypeError("cannot provide both step and domain")
        if domain is not None:
            if isinstance(domain, (int, float)):
                domain = ((0, domain),) * len(shape)
            elif len(domain) != len(shape):
                raise ValueError(f"length of domain does not match ndim: {len(domain)} != {len(shape)}")
            domain = tuple((float(lo), float(hi)) for lo, hi in domain)
        else:
            if step is None:
                step = 1
            if isinstance(step, numbers.Number):
                step = (step,) * len(shape)
            elif len(step) != len(shape):
                raise ValueError(f"length of step does not match ndim: {len(step)} != {len(shape)}")
            domain = tuple((0.0, float(s * n)) for s, n in zip(step, shape))
        object.__setattr__(self, "domain", domain)
        object.__setattr__(self, "step", tuple((hi - lo) / n for (lo, hi), n in zip(domain, shape)))

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def cell_center(self):
        return self.ndim * (0.5,)

    @property
    def cell_faces(self):
        d = self.ndim
        offsets = (np.eye(d) + np.ones([d, d])) / 2.0
        return tuple(tuple(float(o) for o in off) for off in offsets)

    def stagger(self, v):
        return tuple(GridArray(u, o, self) for u, o in zip(v, self.cell_faces))

    def center(self, v):
        return GridArray(v, self.cell_center, self)

    def axes(self, offset=None):
        """float32 arithmetic as JAX would do it: lower + (arange + offset) * step."""
        if offset is None:
            offset = self.cell_center
        if len(offset) != self.ndim:
            raise ValueError(f"unexpected offset length: {len(offset)} vs {self.ndim}")
        f = np.float32
        return tuple(torch.from_numpy(f(lo) + (np.arange(n, dtype=f) + f(o)) * f(h))
                     for (lo, _), o, n, h in zip(self.domain, offset, self.shape, self.step))

    def mesh(self, offset=None):
        return tuple(torch.meshgrid(*self.axes(offset), indexing="ij"))

    def eval_on_mesh(self, fn: Callable, offset=None) -> "GridArray":
        if offset is None:
            offset = self.cell_center
        return GridArray(fn(*self.mesh(offset)), offset, self)


@dataclasses.dataclass
class GridArray:
    """grids.py:107-177: data + offset + grid, with offset-checked arithmetic."""

    data: Any
    offset: Tuple[float, ...]
    grid: Grid

    def __post_init__(self):
        self.data = _to_tensor(self.data)
        self.offset = tuple(float(o) for o in self.offset)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def shape(self):
        return tuple(self.data.shape)

    def to(self, device) -> "GridArray":
        return GridArray(self.data.to(device), self.offset, self.grid)

    def _binary(self, other, op):
        if isinstance(other, GridArray):
            if other.offset != self.offset:
                raise InconsistentOffsetError(f"arrays do not have a unique offset: {self.offset} vs {other.offset}")
            if other.grid != self.grid:
                raise InconsistentGridError("arrays do not have a unique grid")
            other = other.data
        return GridArray(op(self.data, other), self.offset, self.grid)

    def __add__(self, o): return self._binary(o, operator.add)
    def __radd__(self, o): return self._binary(o, lambda a, b: b + a)
    def __sub__(self, o): return self._binary(o, operator.sub)
    def __rsub__(self, o): return self._binary(o, lambda a, b: b - a)
    def __mul__(self, o): return self._binary(o, operator.mul)
    def __rmul__(self, o): return self._binary(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._binary(o, operator.truediv)
    def __neg__(self): return GridArray(-self.data, self.offset, self.grid)


GridArrayVector = Tuple[GridArray, ...]


@dataclasses.dataclass
class GridVariable:
    """grids.py:331-458: a GridArray with its boundary conditions."""

    array: GridArray
    bc: Any

    def __post_init__(self):
        if not isinstance(self.array, GridArray):
            raise ValueError(f"Expected array type to be GridArray, got {type(self.array)}")
        if len(self.bc.types) != self.array.grid.ndim:
            raise ValueError("Incompatible dimension between grid and bc")

    @property
    def data(self): return self.array.data
    @property
    def offset(self): return self.array.offset
    @property
    def grid(self): return self.array.grid
    @property
    def dtype(self): return self.array.dtype
    @property
    def shape(self): return self.array.shape

    def shift(self, offset: int, axis: int) -> GridArray:
        return self.bc.shift(self.array, offset, axis)

    def trim_boundary(self) -> GridArray:
        return self.bc.trim_boundary(self.array)

    def impose_bc(self) -> "GridVariable":
        return self.bc.impose_bc(self.array)

    def to(self, device) -> "GridVariable":
        return GridVariable(self.array.to(device), self.bc)


GridVariableVector = Tuple[GridVariable, ...]


def averaged_offset(*arrays):
    return tuple(float(x) for x in np.mean([a.offset for a in arrays], axis=0))


def control_volume_offsets(c):
    """grids.py:511-517."""
    return tuple(tuple(o + 0.5 if i == j else o for i, o in enumerate(c.offset)) for j in range(len(c.offset)))


def consistent_grid(*arrays) -> Grid:
    grids = {a.grid for a in arrays}
    if len(grids) != 1:
        raise InconsistentGridError(f"arrays do not have a unique grid: {grids}")
    return grids.pop()


def applied(func):
    """grids.py:464-497: lift an array function to GridArrays sharing one offset."""

    def wrapper(*args, **kwargs):
        ga = [a for a in args if isinstance(a, GridArray)] + [v for v in kwargs.values() if isinstance(v, GridArray)]
        offsets = {a.offset for a in ga}
        if len(offsets) != 1:
            raise InconsistentOffsetError(f"arrays do not have a unique offset: {offsets}")
        raw = [a.data if isinstance(a, GridArray) else a for a in args]
        rawk = {k: (v.data if isinstance(v, GridArray) else v) for k, v in kwargs.items()}
        return GridArray(func(*raw, **rawk), ga[0].offset, consistent_grid(*ga))

    return wrapper
