import numpy as _np

Array = _np.ndarray
f32 = _np.float32
f64 = _np.float32  # x64 disabled


def static_cast(*xs):
    return tuple(_np.float32(x) for x in xs)
