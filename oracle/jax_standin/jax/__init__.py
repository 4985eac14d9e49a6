"""NumPy-backed STAND-IN for the few `jax` entry points that hashimmg/jax_IB's own modules use.  TEST INFRASTRUCTURE ONLY.

JAX cannot be installed in the build image, so the reference cannot run as published.  This package lets the
reference's UNMODIFIED source files (imported from /root/reference by tests/golden/make_reference_golden.py, never
copied) execute on NumPy so that the oracle can be pinned against the reference's own code instead of only against
mathematical identities.  It is not JAX: no tracing, no XLA.  What it reproduces is the arithmetic contract the
oracle is held to -- float32 / int32 everywhere (x64 disabled), weakly typed Python scalars -- through an ndarray
subclass that rounds every result back to 32 bits (computing a basic operation in float64 on float32 operands and
rounding once more to float32 is bit-identical to the float32 operation: 53 >= 2 * 24 + 2).  Transcendental functions
(exp, sin, cos) are NumPy's: last-ulp differences from XLA's are possible, which is why the pinned comparisons use a
1e-6 bar, not bit equality.  jax.jacrev is a complex-step derivative (exact to rounding for the analytic kinematic
laws it is applied to).

Nothing under jax_ib_b200/ imports this package; only the fixture generator puts its parent directory on sys.path.
"""
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import core, lax, random, tree_util  # noqa: F401
from . import interpreters, scipy  # noqa: F401
from .numpy import Arr, _wrap

Array = Arr
np = numpy  # the reference's grids.py spells it `jax.np` in one annotation


def named_call(f=None, *, name=None):
    return f


def device_get(x):
    return x


def device_count():
    return 1


def jit(f, *a, **k):
    return f


def vmap(f, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax]
                break
        outs = []
        for k in range(n):
            sl = [a if ax is None else _wrap(_np.take(_np.asarray(a), k, axis=ax)) for a, ax in zip(args, axes)]
            outs.append(f(*sl))
        if isinstance(outs[0], (tuple, list)):
            return type(outs[0])(_wrap(_np.stack([_np.asarray(o[i]) for o in outs], axis=out_axes)) for i in range(len(outs[0])))
        return _wrap(_np.stack([_np.asarray(o) for o in outs], axis=out_axes))
    return mapped


def pmap(f, *a, **k):
    return vmap(f)


def jacrev(f, argnums=0):
    """Complex-step derivative of a scalar-argument function (the only use in the reference: d/dt of the kinematic
    laws, IBM_Force.py:101-102, particle_motion.py)."""
    def df(t, *rest):
        h = 1e-20
        z = f(complex(float(t), h), *rest)
        return _wrap((_np.imag(_np.asarray(z, dtype=_np.complex128)) / h).astype(_np.float32))
    return df


grad = jacrev


def tree_map(fn, tree, *rest):
    return tree_util.tree_map(fn, tree, *rest)


def tree_flatten(tree):
    return tree_util.tree_flatten(tree)


def tree_unflatten(treedef, leaves):
    return tree_util.tree_unflatten(treedef, leaves)
