"""Builds reference-shaped state (All_Variables) and a FusedStepper from a configs.py description --
the equivalent of the set-up cells of the reference's notebooks (Flapping_Demo.ipynb cells 3/5/8)."""
from __future__ import annotations

import numpy as np
import torch

from . import boundaries, configs, engine, grids, kinematics, ops, particle_class

DISPLACEMENT = {"harmonic": kinematics.displacement, "harmonic_multi": kinematics.displacement_multiple, "fourier": kinematics.Displacement_Foil_Fourier_Dotted_Mutliple}
ROTATION = {
    "harmonic": kinematics.rotation,
    "harmonic_multi": kinematics.rotation_multiple,
    "fourier": kinematics.rotation_Foil_Fourier_Dotted_Mutliple,
    "fourier_normalized": kinematics.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED,
}


def _zero_fn(t):
    return 0.0


def make_particles(cfg):
    b = cfg["bodies"]
    if b is None:
        return None
    kw = {"ntheta": b["ntheta"]} if "ntheta" in b else {}
    fn = configs.SHAPES[b["shape"]]
    shape_fn = lambda geometry_param, theta: fn(geometry_param, **kw)
    return particle_class.particle(
        np.asarray(b["centers"], dtype=np.float32), np.asarray(b["geometry"], dtype=np.float32),
        b["displacement_param"], b["rotation_param"], particle_class.Grid1d(2, domain=(0, 2 * np.pi)), shape_fn,
        DISPLACEMENT[b["displacement"]], ROTATION[b["rotation"]])


def make_all_variables(cfg, device="cpu", fields=None):
    grid = grids.Grid(cfg["shape"], domain=cfg["domain"])
    fns = [_zero_fn] * 4
    wall_u, wall_v = cfg["wall_u"], cfg["wall_v"]
    vx_bc = ((0.0, 0.0), tuple(wall_u))
    vy_bc = ((0.0, 0.0), tuple(wall_v))
    if cfg["bc"] == "periodic":
        mk = boundaries.new_periodic_boundary_conditions
    else:
        mk = boundaries.Moving_wall_boundary_conditions
        fns_u = [_zero_fn, _zero_fn, lambda t: wall_u[0], lambda t: wall_u[1]]
        fns_v = [_zero_fn, _zero_fn, lambda t: wall_v[0], lambda t: wall_v[1]]
    if cfg["bc"] == "periodic":
        bcs = (mk(ndim=2, bc_vals=vx_bc, bc_fn=fns, time_stamp=np.float32(0.0)),
               mk(ndim=2, bc_vals=vy_bc, bc_fn=fns, time_stamp=np.float32(0.0)))
    else:
        bcs = (mk(ndim=2, bc_vals=vx_bc, bc_fn=fns_u, time_stamp=np.float32(0.0)),
               mk(ndim=2, bc_vals=vy_bc, bc_fn=fns_v, time_stamp=np.float32(0.0)))
    u, v, p = fields if fields is not None else configs.initial_fields(cfg)
    tens = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in (u, v, p)]
    vel = tuple(grids.GridVariable(grids.GridArray(t, off, grid), bc) for t, off, bc in zip(tens[:2], grid.cell_faces, bcs))
    pres = grids.GridVariable(grids.GridArray(tens[2], grid.cell_center, grid), boundaries.get_pressure_bc_from_velocity(vel))
    return particle_class.All_Variables(make_particles(cfg), vel, pres, [0], 0, [0])


def make_spec(cfg, batch=1, ib_window=6) -> ops.GridSpec:
    nx, ny = cfg["shape"]
    (x0, x1), (y0, y1) = cfg["domain"]
    return ops.GridSpec(nx=nx, ny=ny, lower=(x0, y0), step=((x1 - x0) / nx, (y1 - y0) / ny), dt=cfg["dt"],
                        nu=cfg["viscosity"] / cfg["density"], density=cfg["density"], bc=cfg["bc"],
                        convect=cfg["convect"], u_wall=tuple(cfg["wall_u"]), v_wall=tuple(cfg["wall_v"]),
                        force=tuple(cfg["force"]), ib_window=ib_window, batch=batch)


def make_stepper(cfg, batch=1, ib_window=6, device="cuda") -> engine.FusedStepper:
    return engine.FusedStepper(make_spec(cfg, batch, ib_window), make_particles(cfg), device=device)
