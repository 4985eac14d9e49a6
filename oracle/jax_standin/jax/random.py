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


NUMPY_UNIFORM_FOR_INPUTS = False


def uniform(key, shape=(), dtype=_np.float32, minval=0.0, maxval=1.0):
    """NOT JAX's stream.  Only for notebook cells that draw INPUT data (tracer start positions) whose values are stored
    with the fixture or do not influence what is compared; the generator has to opt in through the module flag."""
    if not NUMPY_UNIFORM_FOR_INPUTS:
        raise NotImplementedError("threefry is not restated")
    rng = _np.random.default_rng(int(_np.asarray(key).ravel()[-1]))
    return (minval + (maxval - minval) * rng.random(shape)).astype(_np.float32)
