"""The full time step.  ORACLE (tests only).

Restates:
  * jax_ib/base/particle_class.py:144-163    All_Variables
  * jax_ib/base/time_stepping.py:109-257     navier_stokes_rk_updated with the forward-Euler tableau
                                             (forward_euler_updated :378-387): the IB step
  * jax_ib/base/time_stepping.py:259-376     navier_stokes_rk_penalty / forward_euler_penalty
  * jax_ib/base/boundaries.py:519-542        update_BC (the clock: time_stamp += dt)
  * jax_ib/base/particle_motion.py:6-66      Update_particle_position_Multiple[_and_MD_Step]
  * jax_cfd.base.funcutils.repeated/trajectory (un-vendored; lax.scan loops)

Step order (time_stepping.py:187-253):
    k = F(u);  dP = grad(p);  u* = u + dt k - dP;  f = IBM(u*);  u** = u* + dt f;
    (u, p) = project(u**, p);  t += dt;  centres += dt Xdot(t_new);  Step_count += 1;  [tracer step]
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Optional, Sequence

import numpy as np

from .field import BC, Var
from . import ib, poisson, stencils


@dataclasses.dataclass
class State:
    """All_Variables (particle_class.py:144-163)."""

    particles: Optional[ib.Bodies]
    velocity: tuple
    pressure: Var
    Drag: Any = 0
    Step_count: int = 0
    MD_var: Any = None


def update_bc(state: State, dt) -> State:
    """boundaries.py:519-542: ts = time_stamp + dt (in the clock's dtype); wall values bc_fn[k](ts)."""
    v = state.velocity
    dtype = v[0].data.dtype.type
    ts = dtype(v[0].bc.time_stamp) + dtype(dt)
    new_v = []
    for u in v:
        fns = u.bc.boundary_fn
        values = u.bc.values if fns is None else ((fns[0](ts), fns[1](ts)), (fns[2](ts), fns[3](ts)))
        new_v.append(u.with_data(u.data, bc=BC(u.bc.types, values, ts, fns)))
    return dataclasses.replace(state, velocity=tuple(new_v))


def update_positions(state: State, dt, md_step: Optional[Callable] = None) -> State:
    """particle_motion.py:38-66 (and :6-35 with the tracer sub-step): centres += dt * Xdot(t),
    t = velocity[0].bc.time_stamp (already advanced by update_bc); Step_count += 1."""
    bodies = state.particles
    out = dataclasses.replace(state, Step_count=state.Step_count + 1)
    if bodies is not None:
        dtype = state.velocity[0].data.dtype.type
        t = state.velocity[0].bc.time_stamp
        U0 = np.asarray(bodies.displacement[1](list(bodies.displacement_param), t, dtype), dtype=dtype).reshape(2, -1)
        c = np.asarray(bodies.centers, dtype=dtype)
        new_c = np.stack([c[:, 0] + dtype(dt) * U0[0], c[:, 1] + dtype(dt) * U0[1]], axis=1)
        out = dataclasses.replace(out, particles=bodies.replace_centers(new_c))
    if md_step is not None:
        out = dataclasses.replace(out, MD_var=md_step(state))
    return out


def make_step(
    dt: float,
    nu: float,
    density: float = 1.0,
    convect: str = "upwind",
    pressure_solve="periodic",
    forcing: Optional[Callable] = None,
    ib_window: Optional[int] = None,
    md_step: Optional[Callable] = None,
    penalty: bool = False,
    ib_threads: int = 0,
):
    """equations.py:194-242 semi_implicit_navier_stokes_timeBC (penalty=False) or :245-280
    semi_implicit_navier_stokes_penalty (penalty=True), with the forward-Euler stepper."""

    def step(state: State) -> State:
        v = state.velocity
        d = v[0].data.dtype.type
        k = stencils.explicit_terms(v, nu, dt, convect, forcing, density)  # time_stepping.py:191
        if penalty:  # time_stepping.py:345: no stored-pressure gradient, no IB, no body update
            u_star = tuple(u.data + d(dt) * ki for u, ki in zip(v, k))
            vs = tuple(u.with_data(x) for u, x in zip(v, u_star))
            new_v, new_p = poisson.projection_and_update_pressure(vs, state.pressure, pressure_solve)
            return update_bc(dataclasses.replace(state, velocity=new_v, pressure=new_p), dt)
        dP = tuple(stencils.forward_difference(state.pressure, a) for a in range(2))  # :192
        u_star = tuple(u.data + d(dt) * ki - g for u, ki, g in zip(v, k, dP))  # :208
        vs = tuple(u.with_data(x) for u, x in zip(v, u_star))
        if state.particles is not None:
            if ib_threads > 0 and ib_window is None:  # the reference's marker sharding on host threads (bench reference arm)
                force = ib.ibm_force_threads(vs, state.particles, dt, ib_threads)
            else:
                force = ib.ibm_force(vs, state.particles, dt, ib_window)  # :210
            vs = tuple(u.with_data(u.data + d(dt) * f) for u, f in zip(vs, force))  # :223
        new_v, new_p = poisson.projection_and_update_pressure(vs, state.pressure, pressure_solve)  # :246
        out = update_bc(dataclasses.replace(state, velocity=new_v, pressure=new_p), dt)  # :248
        return update_positions(out, dt, md_step)  # :253

    return step


def repeated(step: Callable, steps: int) -> Callable:
    """jax_cfd.base.funcutils.repeated."""

    def run(x):
        for _ in range(steps):
            x = step(x)
        return x

    return run


def trajectory(step: Callable, steps: int, post_process: Callable = lambda x: x, start_with_input: bool = False):
    """jax_cfd.base.funcutils.trajectory: returns (final, [snapshots])."""

    def run(x):
        snaps = []
        for _ in range(steps):
            nxt = step(x)
            snaps.append(post_process(x if start_with_input else nxt))
            x = nxt
        return x, snaps

    return run
