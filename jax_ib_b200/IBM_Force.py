"""Direct-forcing immersed boundary with the reference's names (jax_ib/base/IBM_Force.py:20-115),
evaluated by the CUDA kernels of csrc/ib.cu: host kinematics -> warp-per-marker Gaussian interpolation
-> deterministic tile-gather spreading."""
from __future__ import annotations

import numpy as np
import torch

from . import _probe, engine, grids, kinematics, ops


def calc_IBM_force_NEW_MULTIPLE(all_variables, discrete_fn, surface_fn, dt):
    """IBM_Force.py:109-115: returns (GridVariable f_u, GridVariable f_v), the spread force density."""
    dtag = _probe.try_probe(discrete_fn, _probe.Probe(), 0.0, 1.0)
    stag = _probe.try_probe(surface_fn, _probe.Probe(), _probe.Probe(), _probe.Probe())
    ok = isinstance(dtag, _probe.Tag) and dtag.value == "gaussian" and isinstance(stag, _probe.Tag) and stag.value == "gaussian"
    if not ok:
        raise NotImplementedError("calc_IBM_force_NEW_MULTIPLE: the CUDA path implements the Gaussian delta "
                                  "(convolution_functions.delta_approx_logistjax) with new_surf_fn only")
    window = stag.extra or 6
    if _probe.is_probe(all_variables):
        return _probe.Tag("ibm", "gaussian", window)
    v = all_variables.velocity
    grid = grids.consistent_grid(*v)
    particles = all_variables.particles
    spec = engine.grid_spec_from(grid, dt, None, 1.0, "periodic", "none", ib_window=window)
    bodies = engine.get_bodies(particles, v[0].data.device)
    table, _, _ = kinematics.body_schedule(particles, v[0].bc.time_stamp, dt, 1)
    dev = engine.cuda_device(v[0].data.device)
    state = torch.from_numpy(table[0]).to(dev)
    # spread into zero fields: u** - u* = dt * f  =>  f = result / dt
    fu = torch.zeros(grid.shape, dtype=torch.float32, device=dev)
    fv = torch.zeros(grid.shape, dtype=torch.float32, device=dev)
    us, vs = engine.to_device(v[0].data).clone(), engine.to_device(v[1].data).clone()
    mk = ops.ib_force(spec, bodies, state, us, vs)  # fills marker arrays and applies u* += dt f
    for comp, (out, off) in enumerate(((fu, v[0].offset), (fv, v[1].offset))):
        ops.ib_spread(spec, off, mk[3 + comp, 0], mk[0, 0], mk[1, 0], mk[2, 0], out=out)
    return tuple(grids.GridVariable(grids.GridArray(f, x.offset, grid), x.bc) for f, x in zip((fu, fv), v))


def _scalar(x) -> np.float32:
    return np.float32(x.detach().cpu().reshape(-1)[0] if isinstance(x, torch.Tensor) else np.asarray(x).reshape(-1)[0])


def IBM_force_GENERAL(field, Xi, particle_center, geom_param, Grid_p, shape_fn, discrete_fn, surface_fn, dx_dt, domega_dt,
                      rotation, dt):
    """IBM_Force.py:20-87 for ONE body and ONE velocity component, with the caller's kinematic callables: markers
    rotated and translated on the host in float32, `surface_fn` (the interpolation kernel) for the fluid velocity at
    the markers, (U_body - U) / dt, arc length, the spreading kernel.  Returns the force-density array on the device."""
    tag = _probe.try_probe(discrete_fn, _probe.Probe(), 0.0, 1.0)
    if not (isinstance(tag, _probe.Tag) and tag.value == "gaussian"):
        raise NotImplementedError("IBM_force_GENERAL: the CUDA path implements the Gaussian delta only")
    grid = field.grid
    t = field.bc.time_stamp
    x0, y0 = (np.asarray(a, dtype=np.float32) for a in shape_fn(geom_param, Grid_p))
    theta = _scalar(rotation(t))
    c, s = np.cos(theta), np.sin(theta)
    xc, yc = (_scalar(v) for v in particle_center)
    xp = x0 * c - y0 * s + xc                                                    # :33-34
    yp = x0 * s + y0 * c + yc
    U = surface_fn(field, xp, yp)                                                # :37
    r = -(yp - yc) if Xi == 0 else (xp - xc)                                     # :39-42
    U0 = dx_dt(t)
    U0 = U0.detach().cpu().numpy() if isinstance(U0, torch.Tensor) else np.asarray(U0)
    UP = np.float32(U0.reshape(2, -1)[Xi, 0]) + _scalar(domega_dt(t)) * r        # :44-47
    dev = U.device
    force = (torch.from_numpy(UP.astype(np.float32)).to(dev) - U) / np.float32(dt)  # :50
    dS = np.sqrt((np.roll(xp, -1) - xp) ** 2 + (np.roll(yp, -1) - yp) ** 2).astype(np.float32)  # :59-63
    window = getattr(_probe.try_probe(surface_fn, _probe.Probe(), _probe.Probe(), _probe.Probe()), "extra", None) or 6
    spec = engine.grid_spec_from(grid, dt, None, 1.0, "periodic", "none", ib_window=window)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    return ops.ib_spread(spec, field.offset, force, to(xp), to(yp), to(dS))      # :66-87


def IBM_Multiple_NEW(field, Xi, particles, discrete_fn, surface_fn, dt):
    """IBM_Force.py:89-106: the force density of ALL bodies on one velocity component, a GridArray at field.offset.
    The per-body rates come from the host's body table (closed forms of the library laws, autodiff for user laws)."""
    Grid_p = particles.generate_grid()
    table, _, _ = kinematics.body_schedule(particles, field.bc.time_stamp, dt, 1)
    total = None
    for i in range(len(particles.particle_center)):
        row = table[0, i]
        rot = lambda t, i=i: particles.Rotation_EQ([particles.rotation_param[i]], t)
        f = IBM_force_GENERAL(field, Xi, particles.particle_center[i], particles.geometry_param[i], Grid_p, particles.shape,
                              discrete_fn, surface_fn, lambda t, row=row: row[4:6], lambda t, row=row: row[6], rot, dt)
        total = f if total is None else total + f
    return grids.GridArray(total, field.offset, field.grid)


def integrate_trapz(integrand, dx, dy):
    """IBM_Force.py:5-6: trapezoid rule along the last axis with spacing dx, then along the remaining one with dy."""
    return torch.trapezoid(torch.trapezoid(integrand, dx=dx), dx=dy)


def Integrate_Field_Fluid_Domain(field):
    """IBM_Force.py:9-18."""
    return integrate_trapz(engine.to_device(field.data) if torch.cuda.is_available() else field.data, field.grid.step[0], field.grid.step[1])
