"""User-level finite differences with the reference's names (jax_ib/base/finite_differences.py:46-196),
written with torch ops over GridVariable.shift.  These are for post-processing and for user plug-ins
(e.g. the default `Pressure_Grad=finite_differences.forward_difference`); inside the fused step the
same stencils run inside the CUDA kernels."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import grids

GridArray = grids.GridArray
GridVariable = grids.GridVariable


def stencil_sum(*arrays: GridArray) -> GridArray:
    """finite_differences.py:46-56."""
    offset = grids.averaged_offset(*arrays)
    result = sum(a.data for a in arrays)
    return GridArray(result, offset, grids.consistent_grid(*arrays))


def _axes(u, axis):
    return range(u.grid.ndim) if axis is None else axis


def central_difference(u: GridVariable, axis=None):
    """finite_differences.py:77-84."""
    axis = _axes(u, axis)
    if not isinstance(axis, int):
        return tuple(central_difference(u, a) for a in axis)
    diff = stencil_sum(u.shift(+1, axis), -u.shift(-1, axis))
    return diff / (2 * u.grid.step[axis])


def backward_difference(u: GridVariable, axis=None):
    """finite_differences.py:97-104."""
    axis = _axes(u, axis)
    if not isinstance(axis, int):
        return tuple(backward_difference(u, a) for a in axis)
    diff = stencil_sum(u.array, -u.shift(-1, axis))
    return diff / u.grid.step[axis]


def forward_difference(u: GridVariable, axis=None):
    """finite_differences.py:119-126."""
    axis = _axes(u, axis)
    if not isinstance(axis, int):
        return tuple(forward_difference(u, a) for a in axis)
    diff = stencil_sum(u.shift(+1, axis), -u.array)
    return diff / u.grid.step[axis]


def laplacian(u: GridVariable) -> GridArray:
    """finite_differences.py:129-136."""
    scales = np.square(1 / np.array(u.grid.step, dtype=np.float32))
    result = -2 * u.array * float(np.sum(scales))
    for axis in range(u.grid.ndim):
        result = result + stencil_sum(u.shift(-1, axis), u.shift(+1, axis)) * float(scales[axis])
    return result


def divergence(v: Sequence[GridVariable]) -> GridArray:
    """finite_differences.py:139-146."""
    grid = grids.consistent_grid(*v)
    if len(v) != grid.ndim:
        raise ValueError(f"The length of `v` must be equal to `grid.ndim`. Expected length {grid.ndim}; got {len(v)}.")
    diffs = [backward_difference(u, axis) for axis, u in enumerate(v)]
    return sum(diffs[1:], diffs[0])


def curl_2d(v: Sequence[GridVariable]) -> GridArray:
    """finite_differences.py:189-196."""
    if len(v) != 2:
        raise ValueError(f"Length of `v` is not 2: {len(v)}")
    return forward_difference(v[1], axis=0) - forward_difference(v[0], axis=1)
