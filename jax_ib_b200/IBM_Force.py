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
