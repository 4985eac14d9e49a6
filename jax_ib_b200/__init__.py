"""jax_ib_b200: B200-native (sm_100a CUDA) implementation of the per-timestep hot path of
hashimmg/jax_IB behind the reference's own pluggable-callable API.  See DESIGN.md / INTEGRATION.md.

    import jax_ib_b200 as ib
    step_fn = ib.equations.semi_implicit_navier_stokes_timeBC(density, viscosity, dt, grid, convect=..., ...)

Importing the package does not need a GPU; running any operator does (there is no CPU fallback).
"""
from . import (IBM_Force, MD, advection, boundaries, configs, convolution_functions, demo, diffusion, engine,  # noqa: F401
               equations, finite_differences, funcutils, grids, kinematics, ops, particle_class, particle_motion,
               penalty, pressure, time_stepping)

base = __import__(__name__)  # `import jax_ib_b200.base as ib` style access: ib.equations, ib.grids, ...
__version__ = "0.1"
