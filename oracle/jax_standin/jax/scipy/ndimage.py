"""jax.scipy.ndimage stand-in: map_coordinates(order=1), restating JAX's published algorithm rather than SciPy's
(SciPy's mode='constant' returns cval for any point outside the array; JAX validates TAP BY TAP, so a point half a
cell outside still picks up its inside neighbours -- SciPy calls that 'grid-constant'):

    per axis: lower = floor(x), taps (lower, 1 - (x - lower)) and (lower + 1, x - lower); a tap is valid when its
    index lies in [0, size); indices are clipped for the gather; over the product of taps, in order,
    contribution = where(all valid, input[idx], cval) * prod(weights); result = left-to-right sum.  float32.
"""
import functools
import itertools
import operator

import numpy as _np

from ..numpy import _plain, _wrap


def map_coordinates(input, coordinates, order, mode="constant", cval=0.0):
    if order != 1 or mode != "constant":
        raise NotImplementedError("only order=1, mode='constant' is used by the reference (CFD_MD_coupling.py:5)")
    a = _np.asarray(_plain(input))
    t = a.dtype.type
    per_axis = []
    for c, size in zip(coordinates, a.shape):
        c = _np.asarray(_plain(c)).astype(a.dtype)
        lower = _np.floor(c)
        wu = (c - lower).astype(a.dtype)
        wl = (t(1) - wu).astype(a.dtype)
        idx = lower.astype(_np.int32)
        per_axis.append([(_np.clip(i, 0, size - 1), (i >= 0) & (i < size), w) for i, w in ((idx, wl), (idx + 1, wu))])
    outs = []
    for items in itertools.product(*per_axis):
        idx, valid, w = zip(*items)
        contrib = _np.where(functools.reduce(operator.and_, valid), a[idx], t(cval))
        outs.append(functools.reduce(operator.mul, w) * contrib)
    return _wrap(functools.reduce(operator.add, outs).astype(a.dtype))
