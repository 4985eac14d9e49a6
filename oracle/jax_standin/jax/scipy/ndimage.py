"""jax.scipy.ndimage stand-in: map_coordinates(order=1, mode='constant') = bilinear with zeros outside (the JAX default)."""
import numpy as _np
import scipy.ndimage as _ndi

from ..numpy import _plain, _wrap


def map_coordinates(input, coordinates, order, mode="constant", cval=0.0):
    return _wrap(_ndi.map_coordinates(_np.asarray(_plain(input), dtype=_np.float64), [_np.asarray(_plain(c), dtype=_np.float64) for c in coordinates],
                                      order=order, mode="constant" if mode == "constant" else mode, cval=cval).astype(_np.float32))
