from typing import Callable

import jax.numpy as jnp

Box = object
ShiftFn = Callable
DisplacementFn = Callable


def free():
    return (lambda a, b, **k: a - b), (lambda R, dR, **k: R + dR)


def periodic(side, wrapped=True):
    def displacement(a, b, **k):
        d = a - b
        return d - side * jnp.round(d / side)

    def shift(R, dR, **k):
        return jnp.mod(R + dR, side) if wrapped else R + dR
    return displacement, shift


def canonicalize_displacement_or_metric(fn):
    return fn
