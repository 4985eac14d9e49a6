"""jax.random stand-in for the ZERO-TEMPERATURE Brownian step only (MD/simulate.py:88-94 draws a normal deviate and
multiplies it by sqrt(2 kT dt nu) = 0).  JAX's threefry generator is not restated: `normal` refuses to hand out
numbers unless told the temperature is zero through the module flag the fixture generator sets."""
import numpy as _np

ZERO_TEMPERATURE_ONLY = False


def PRNGKey(seed):
    return _np.array([0, seed], dtype=_np.uint32)


def split(key, num=2):
    return tuple(key for _ in range(num))


def normal(key, shape, dtype=_np.float32):
    if not ZERO_TEMPERATURE_ONLY:
        raise NotImplementedError("threefry / normal deviates are not restated; only kT = 0 runs are generated")
    return _np.zeros(shape, dtype=_np.float32)


def uniform(*a, **k):
    raise NotImplementedError("threefry is not restated")
