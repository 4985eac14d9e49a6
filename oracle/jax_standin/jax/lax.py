"""jax.lax stand-in: slice_in_dim (boundaries.py), scan / Precision (array_utils.py)."""
import numpy as _np

from .numpy import _plain, _wrap


def slice_in_dim(operand, start_index, limit_index, stride=1, axis=0):
    a = _plain(operand)
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(start_index, limit_index, stride)
    return _wrap(a[tuple(idx)])


def scan(f, init, xs, length=None):
    carry, ys = init, []
    n = length if xs is None else len(xs)
    for k in range(n):
        carry, y = f(carry, None if xs is None else xs[k])
        ys.append(y)
    return carry, _wrap(_np.stack([_np.asarray(y) for y in ys])) if ys and ys[0] is not None else None


class Precision:
    HIGHEST = "highest"
