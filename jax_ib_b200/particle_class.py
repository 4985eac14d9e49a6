"""State containers mirrored from the reference's jax_ib/base/particle_class.py: Grid1d (:21-101),
particle (:106-139) and All_Variables (:144-163).  Plain dataclasses (there is no pytree machinery
to satisfy here); field names and order follow the reference."""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Sequence

import numpy as np


@dataclasses.dataclass(init=False, frozen=True)
class Grid1d:
    """particle_class.py:21-101: theta grid for parametrised shapes (step = span / (n - 1))."""

    shape: int
    step: float
    domain: tuple

    def __init__(self, shape, step=None, domain=None):
        object.__setattr__(self, "shape", shape)
        object.__setattr__(self, "domain", domain)
        object.__setattr__(self, "step", (domain[1] - domain[0]) / (shape - 1))

    @property
    def ndim(self):
        return 1

    @property
    def cell_center(self):
        return (0.5,)

    def axes(self, offset=None):
        return np.float32(self.domain[0]) + np.arange(self.shape, dtype=np.float32) * np.float32(self.step)

    def mesh(self, offset=None):
        return self.axes(offset)


@dataclasses.dataclass
class particle:
    """particle_class.py:106-139."""

    particle_center: Any
    geometry_param: Any
    displacement_param: Any
    rotation_param: Any
    Grid: Grid1d
    shape: Callable
    Displacement_EQ: Callable
    Rotation_EQ: Callable

    def generate_grid(self):
        return self.Grid.mesh()

    def calc_Rtheta(self):
        return self.shape(self.geometry_param, self.Grid)


@dataclasses.dataclass
class All_Variables:
    """particle_class.py:144-163."""

    particles: Any
    velocity: tuple
    pressure: Any
    Drag: Any
    Step_count: Any
    MD_var: Any
