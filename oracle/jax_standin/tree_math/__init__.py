"""tree_math STAND-IN (google/tree-math is a dependency of the reference that is absent here).  TEST INFRASTRUCTURE ONLY.

The three names the reference's equations.py / time_stepping.py use, with the published semantics:

  Vector(tree)      a pytree treated as one vector: + - * / act leaf by leaf, scalars broadcast
  wrap(fun)         `fun` is written on Vectors; the result takes and returns plain pytrees
  unwrap(fun, vector_argnums=0)
                    `fun` is written on pytrees; the result takes a Vector in the named positions and hands the
                    whole result back as one Vector
"""
import functools
import operator

from jax import tree_util as _tu


def _is_vector(x):
    return isinstance(x, Vector)


class Vector:
    def __init__(self, tree):
        self.tree = tree

    def _binary(self, other, op, swap=False):
        if isinstance(other, Vector):
            a, b = (other, self) if swap else (self, other)
            return Vector(_tu.tree_map(op, a.tree, b.tree))
        if swap:
            return Vector(_tu.tree_map(lambda x: op(other, x), self.tree))
        return Vector(_tu.tree_map(lambda x: op(x, other), self.tree))

    def __add__(self, o): return self._binary(o, operator.add)
    def __radd__(self, o): return self._binary(o, operator.add, swap=True)
    def __sub__(self, o): return self._binary(o, operator.sub)
    def __rsub__(self, o): return self._binary(o, operator.sub, swap=True)
    def __mul__(self, o): return self._binary(o, operator.mul)
    def __rmul__(self, o): return self._binary(o, operator.mul, swap=True)
    def __truediv__(self, o): return self._binary(o, operator.truediv)
    def __neg__(self): return Vector(_tu.tree_map(operator.neg, self.tree))


def _argnums(vector_argnums, nargs):
    if vector_argnums is None:
        return tuple(range(nargs))
    if isinstance(vector_argnums, int):
        return (vector_argnums,)
    return tuple(vector_argnums)


def _strip(out):
    """Vector -> its tree, also inside tuples / lists of Vectors."""
    if _is_vector(out):
        return out.tree
    if isinstance(out, (tuple, list)):
        return type(out)(_strip(o) for o in out)
    return out


def wrap(fun, vector_argnums=0):
    @functools.wraps(fun)
    def wrapper(*args, **kwargs):
        idx = _argnums(vector_argnums, len(args))
        args = tuple(Vector(a) if k in idx else a for k, a in enumerate(args))
        return _strip(fun(*args, **kwargs))
    return wrapper


def unwrap(fun, vector_argnums=0, out_vectors=True):
    @functools.wraps(fun)
    def wrapper(*args, **kwargs):
        idx = _argnums(vector_argnums, len(args))
        args = tuple(a.tree if (k in idx and _is_vector(a)) else a for k, a in enumerate(args))
        out = fun(*args, **kwargs)
        return Vector(out) if out_vectors else out
    return wrapper
