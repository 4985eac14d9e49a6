"""Solver assembly with the reference's public API (jax_ib/base/equations.py:194-280):

    step_fn = semi_implicit_navier_stokes_timeBC(density, viscosity, dt, grid, convect=..., diffuse=...,
                  pressure_solve=..., forcing=..., time_stepper=..., IBM_forcing=..., Updating_Position=...,
                  Pressure_Grad=..., Drag_fn=...)
    all_variables = step_fn(all_variables)

The plug-ins are probed once (jax_ib_b200/_probe.py).  When they are this package's own library
functions (the way every notebook of the reference wires them), `step_fn` is ONE fused CUDA pipeline
(engine.FusedStepper: predict -> IB -> FFT projection, device resident) and additionally exposes
`step_fn.advance(x, n)` which `funcutils.repeated` uses to run a whole inner loop per host round trip.
Opaque user plug-ins are honoured by executing the reference's step order (time_stepping.py:187-253)
stage by stage on the GPU with the user's callables.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import numpy as np
import torch

from . import _probe, advection, boundaries, diffusion, engine, finite_differences, grids, pressure, time_stepping

GridArray = grids.GridArray
GridVariable = grids.GridVariable


def constant_forcing(force_vector):
    """forcing(v) = constant body force per component (the Taylor demo's `pressure_gradient_forcing`,
    taylor_dispersion_demo.ipynb cell 10), tagged so that it fuses into the predictor kernel."""
    fx, fy = float(force_vector[0]), float(force_vector[1])

    def forcing(v):
        return tuple(GridArray(torch.full_like(u.data, f), u.offset, u.grid) for f, u in zip((fx, fy), v))

    forcing._jaxib_constant = (fx, fy)
    return forcing


def _recognise(convect, diffuse, pressure_solve, forcing, time_stepper, IBM_forcing, Updating_Position, Pressure_Grad,
               Drag_fn, dt):
    """Returns a dict describing the fused configuration, or None if any plug-in is opaque."""
    cfg = {}
    if convect is None:
        cfg["convect"] = "van_leer_limiters"  # equations.py:55-58
    else:
        tags = _probe.tags_of(_probe.try_probe(convect, _probe.velocity_probe()), "convect")
        if not tags or len({t.value for t in tags}) != 1:
            return None
        cfg["convect"] = tags[0].value
    if diffuse is not diffusion.diffuse:
        t = _probe.try_probe(diffuse, _probe.Probe(), 1.0)
        if not (isinstance(t, _probe.Tag) and t.what == "diffuse"):
            return None
    t = _probe.tags_of(_probe.try_probe(pressure_solve, _probe.velocity_probe(), None), "pressure_solve")
    if not t:
        return None
    cfg["poisson"] = t[0].value
    cfg["force"], cfg["perm"] = (0.0, 0.0), None
    if forcing is not None:
        if hasattr(forcing, "_jaxib_constant"):
            cfg["force"] = forcing._jaxib_constant
        elif hasattr(forcing, "_jaxib_penalty"):
            cfg["force"], cfg["perm"] = forcing._jaxib_penalty
        else:
            return None
    if not isinstance(time_stepper, time_stepping._Stepper):
        return None
    cfg["incremental"] = time_stepper.incremental_pressure
    cfg["window"], cfg["ib"] = 6, False
    if IBM_forcing is not None:
        t = _probe.try_probe(IBM_forcing, _probe.Probe("all"), dt)
        if not (isinstance(t, _probe.Tag) and t.what == "ibm"):
            return None
        cfg["ib"], cfg["window"] = True, t.extra or 6
    cfg["md_step"], cfg["move"] = None, False
    if Updating_Position is not None:
        t = _probe.try_probe(Updating_Position, _probe.Probe("all"), dt)
        if not (isinstance(t, _probe.Tag) and t.what == "position"):
            return None
        cfg["move"], cfg["md_step"] = True, t.extra
    if Pressure_Grad is not finite_differences.forward_difference:
        return None
    if Drag_fn is not None:
        sentinel = _probe.Probe("all")
        if _probe.try_probe(Drag_fn, sentinel, dt) is not sentinel:
            return None  # a real in-loop reduction: needs the stage-by-stage path
    return cfg


class _StepFn:
    def __init__(self, density, viscosity, dt, grid, cfg, general):
        self.density, self.viscosity, self.dt, self.grid, self.cfg = density, viscosity, dt, grid, cfg
        self._general = general
        self._stepper = None
        self.fused = cfg is not None

    def _get_stepper(self, x):
        if self._stepper is None:
            c = self.cfg
            kind = engine.bc_kind_of(x.velocity)
            if c["poisson"] != kind:
                raise ValueError(f"pressure_solve is for {c['poisson']} grids but the velocity BCs are {kind}")
            spec = engine.grid_spec_from(self.grid, self.dt, self.viscosity, self.density, kind, c["convect"], x.velocity,
                                         c["force"], c["window"])
            perm = engine.to_device(c["perm"].data if hasattr(c["perm"], "data") else c["perm"]) if c["perm"] is not None else None
            self._stepper = engine.FusedStepper(spec, x.particles if c["ib"] else None, perm, c["incremental"], c["md_step"],
                                                device=engine.cuda_device(x.velocity[0].data.device))
        return self._stepper

    def advance(self, x, nsteps: int):
        if not self.fused:
            for _ in range(nsteps):
                x = self._general(x)
            return x
        return self._get_stepper(x).advance(x, nsteps, move_bodies=self.cfg["move"])

    def rollout(self, x, inner_steps: int, outer_steps: int, start_with_input: bool = False):
        """trajectory(repeated(self, inner), outer) on the device; None when this step function is not fused (or
        carries a per-step host hook), in which case funcutils falls back to the step-by-step loop."""
        if not self.fused or self.cfg["md_step"] is not None:
            return None
        return self._get_stepper(x).rollout(x, inner_steps, outer_steps, start_with_input, move_bodies=self.cfg["move"])

    def __call__(self, x):
        return self.advance(x, 1)


def _general_step(density, viscosity, dt, convect, diffuse, pressure_solve, forcing, incremental, IBM_forcing,
                  Updating_Position, Pressure_Grad, Drag_fn):
    """time_stepping.py:144-255 with the user's callables, every stage on the GPU."""
    if convect is None:
        convect = lambda v: tuple(advection.advect_van_leer_using_limiters(u, v, dt) for u in v)

    def step(x):
        v = tuple(u.to(engine.cuda_device(u.data.device)) for u in x.velocity)
        x = dataclasses.replace(x, velocity=v, pressure=x.pressure.to(v[0].data.device))
        k = list(convect(v))                                                     # equations.py:70-77
        if viscosity is not None:
            k = [a + diffuse(u, viscosity / density) for a, u in zip(k, v)]
        if forcing is not None:
            k = [a + f / density for a, f in zip(k, forcing(v))]
        u_star = [u.array + dt * a for u, a in zip(v, k)]
        if incremental:
            dP = Pressure_Grad(x.pressure)                                        # time_stepping.py:192
            u_star = [a - g for a, g in zip(u_star, dP)]
        vs = tuple(GridVariable(a, u.bc) for a, u in zip(u_star, v))
        x = dataclasses.replace(x, velocity=vs)
        if IBM_forcing is not None:
            force = IBM_forcing(x, dt)                                            # :210
            if Drag_fn is not None:
                x = dataclasses.replace(x, Drag=Drag_fn(dataclasses.replace(x, velocity=force), dt).Drag)
            vs = tuple(GridVariable(u.array + dt * f.array, u.bc) for u, f in zip(vs, force))  # :223
            x = dataclasses.replace(x, velocity=vs)
        x = pressure.projection_and_update_pressure(x, pressure_solve)            # :246
        x = boundaries.update_BC(x, dt)                                           # :248
        if Updating_Position is not None:
            x = Updating_Position(x, dt)                                          # :253
        return x

    return step


def semi_implicit_navier_stokes_timeBC(
    density: float,
    viscosity: float,
    dt: float,
    grid: grids.Grid,
    convect: Optional[Callable] = None,
    diffuse: Callable = diffusion.diffuse,
    pressure_solve: Callable = pressure.solve_fast_diag,
    forcing: Optional[Callable] = None,
    time_stepper=time_stepping.forward_euler_updated,
    IBM_forcing: Optional[Callable] = None,
    Updating_Position: Optional[Callable] = None,
    Pressure_Grad: Callable = finite_differences.forward_difference,
    Drag_fn: Optional[Callable] = None,
):
    """equations.py:194-242: returns step_fn(All_Variables) -> All_Variables."""
    cfg = _recognise(convect, diffuse, pressure_solve, forcing, time_stepper, IBM_forcing, Updating_Position,
                     Pressure_Grad, Drag_fn, dt)
    incremental = getattr(time_stepper, "incremental_pressure", True)
    general = _general_step(density, viscosity, dt, convect, diffuse, pressure_solve, forcing, incremental, IBM_forcing,
                            Updating_Position, Pressure_Grad, Drag_fn)
    return _StepFn(density, viscosity, dt, grid, cfg, general)


def semi_implicit_navier_stokes_penalty(
    density: float,
    viscosity: float,
    dt: float,
    grid: grids.Grid,
    convect: Optional[Callable] = None,
    diffuse: Callable = diffusion.diffuse,
    pressure_solve: Callable = pressure.solve_fast_diag,
    forcing: Optional[Callable] = None,
    time_stepper=time_stepping.forward_euler_penalty,
):
    """equations.py:245-280: no stored-pressure gradient, no IB stage, no body update."""
    return semi_implicit_navier_stokes_timeBC(density, viscosity, dt, grid, convect, diffuse, pressure_solve, forcing,
                                              time_stepper, None, None, finite_differences.forward_difference, None)
