"""jax.numpy stand-in: NumPy with 32-bit results (see the package docstring)."""
import numpy as _np


def _down(a):
    if isinstance(a, _np.ndarray):
        if a.dtype == _np.float64:
            a = a.astype(_np.float32)
        elif a.dtype == _np.int64:
            a = a.astype(_np.int32)
        elif a.dtype == _np.complex128:
            # complex values only occur inside the complex-step derivative (jax.jacrev stand-in): keep the precision
            pass
        return a.view(Arr)
    if isinstance(a, _np.floating) and a.dtype == _np.float64:
        return _np.float32(a)
    if isinstance(a, _np.integer) and a.dtype == _np.int64:
        return _np.int32(a)
    return a


def _wrap(out):
    if isinstance(out, tuple):
        return tuple(_wrap(o) for o in out)
    if isinstance(out, list):
        return [_wrap(o) for o in out]
    return _down(out)


def _plain(x):
    if isinstance(x, Arr):
        return x.view(_np.ndarray)
    if isinstance(x, (tuple, list)):
        return type(x)(_plain(v) for v in x)
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    return x


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Op:
            def set(self, v):
                out = _np.array(_plain(arr), copy=True)
                out[idx] = _plain(v)
                return _wrap(out)

            def add(self, v):
                out = _np.array(_plain(arr), copy=True)
                out[idx] += _plain(v)
                return _wrap(out)

            def multiply(self, v):
                out = _np.array(_plain(arr), copy=True)
                out[idx] *= _plain(v)
                return _wrap(out)
        return _Op()


class Arr(_np.ndarray):
    """ndarray whose every result is rounded back to float32 / int32 (JAX with x64 disabled)."""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = tuple(_plain(x) for x in inputs)
        if out is not None:
            kwargs["out"] = tuple(_plain(o) for o in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        return _wrap(res)

    def __array_function__(self, func, types, args, kwargs):
        res = func(*_plain(args), **_plain(kwargs))
        return _wrap(res)

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def array(obj, dtype=None, **k):
    return _wrap(_np.array(_plain(obj), dtype=dtype, **k))


def asarray(obj, dtype=None, **k):
    return _wrap(_np.asarray(_plain(obj), dtype=dtype, **k))


def __getattr__(name):
    attr = getattr(_np, name)
    if callable(attr) and not isinstance(attr, type):
        def fn(*a, **k):
            return _wrap(attr(*_plain(a), **_plain(k)))
        fn.__name__ = name
        return fn
    return attr


ndarray = Arr
float32, float64, int32, int64, complex64 = _np.float32, _np.float32, _np.int32, _np.int32, _np.complex64  # x64 disabled
pi, e, inf, nan, newaxis = _np.pi, _np.e, _np.inf, _np.nan, None


def trapz(y, x=None, dx=1.0, axis=-1):
    return _wrap(_np.trapezoid(_plain(y), x=_plain(x), dx=dx, axis=axis))
