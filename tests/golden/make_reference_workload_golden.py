"""Two of the measured workloads advanced by the REFERENCE'S OWN FILES on the NumPy stand-in for jax.  Build container only.

    python tests/golden/make_reference_workload_golden.py ib2048     # ~10 min, ~12 GB peak -> reference_ib2048_3.npz
    python tests/golden/make_reference_workload_golden.py channel    # seconds              -> reference_channel1024x256_5.npz
    python tests/golden/make_reference_workload_golden.py small64    # seconds              -> reference_small64_50.npz

channel: G3's parity case (jax_ib_b200/configs.py::channel: 1024 x 256, periodic x / no-slip y, van Leer advection, body
force, Poiseuille + noise start), five steps, the run tests/golden/channel1024x256_5.npz holds from the oracle.  The
reference's step function always calls its IBM_forcing and Updating_Position plug-ins (time_stepping.py:210,253), so
a body-free run passes a zero force and an identity, as a reference user would.

ib2048: the BENCH WORKLOAD ( 2048 x 2048 periodic, four flapping 298-marker ellipses, upwind, FFT projection -- the
configuration `bench.py` times and BASELINE.json's metric is quoted on) advanced for three steps, wired exactly as
Flapping_Demo.ipynb cell 9 wires the reference.

The reference sums every marker over the WHOLE grid (1192 markers x 4.2 M cells, twice per component and step), which
is why this is three steps and why nothing like it can run inside a test.  Inputs are the product's own workload
description (jax_ib_b200/configs.py::ib_periodic: seeded divergence-free initial flow, body lattice, phases); the
multi-body harmonic laws are user code here as they would be for a reference user (the library's `displacement` /
`rotation` take a single body), written with the library laws' arithmetic.  Stored: four 160 x 160 crops of (u, v, p)
around the bodies, full-field norms and sums, body centres, clock, step count.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
F = np.float32
STEPS = 3
HALF = 80  # crop half-width in cells


def main_ib2048():
    t_start = time.time()
    sys.dont_write_bytecode = True
    sys.path.insert(0, ROOT)
    from jax_ib_b200 import configs
    cfg = configs.ib_periodic(2048, 2)
    u0, v0, p0 = configs.initial_fields(cfg)

    sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_standin"))
    sys.path.insert(0, REF)
    import jax.numpy as jnp
    import jax_cfd.base as cfd
    import jax_ib.base as ib
    from jax_ib.base import IBM_Force, advection, boundaries, convolution_functions, grids, particle_class as pc, particle_motion

    dt, density, viscosity = cfg["dt"], cfg["density"], cfg["viscosity"]
    grid = grids.Grid(cfg["shape"], domain=cfg["domain"])
    zero = lambda t: 0.0
    bc = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=[zero] * 4, time_stamp=0.0)
    v = tuple(grids.GridVariable(grids.GridArray(jnp.asarray(a), off, grid), bc) for a, off in zip((u0, v0), grid.cell_faces))
    p = grids.GridVariable(grids.GridArray(jnp.asarray(p0), grid.cell_center, grid), boundaries.get_pressure_bc_from_velocity(v))

    b = cfg["bodies"]

    def displacement_multi(parameters, t):  # [2, B]: per body [A0/2 cos(2 pi f t), 0]
        A0 = jnp.array([q[0] for q in parameters])
        f = jnp.array([q[1] for q in parameters])
        x = A0 / 2 * jnp.cos(2 * jnp.pi * f * t)
        return jnp.array([x, jnp.zeros_like(x)])

    def rotation_multi(parameters, t):  # [B]: alpha0 + beta sin(2 pi f t + phi)
        a0, beta, f, phi = (jnp.array([q[k] for q in parameters]) for k in range(4))
        return a0 + beta * jnp.sin(2 * jnp.pi * f * t + phi)

    shape_fn = lambda geometry_param, theta: tuple(jnp.asarray(a) for a in configs.ellipse_markers(geometry_param, ntheta=150))
    particles = pc.particle(jnp.asarray(np.asarray(b["centers"], dtype=F)), jnp.asarray(np.asarray(b["geometry"], dtype=F)),
                            [[F(x) for x in q] for q in b["displacement_param"]], [[F(x) for x in q] for q in b["rotation_param"]],
                            pc.Grid1d(2, domain=(0, 2 * np.pi)), shape_fn, displacement_multi, rotation_multi)
    all_variables = pc.All_Variables(particles, v, p, [0], 0, [0])

    discrete_delta = lambda x, x0, w1: convolution_functions.delta_approx_logistjax(x, x0, w1)
    surf_fn = lambda field, xp, yp: convolution_functions.new_surf_fn(field, xp, yp, discrete_delta)
    IBM_forcing = lambda av, dt_: IBM_Force.calc_IBM_force_NEW_MULTIPLE(av, discrete_delta, surf_fn, dt_)
    step_fn = cfd.funcutils.repeated(
        ib.equations.semi_implicit_navier_stokes_timeBC(
            density=density, viscosity=viscosity, dt=dt, grid=grid,
            convect=lambda vv: tuple(advection.advect_upwind(c, vv, dt) for c in vv),
            pressure_solve=ib.pressure.solve_fast_diag, forcing=None, time_stepper=ib.time_stepping.forward_euler_updated,
            IBM_forcing=IBM_forcing, Updating_Position=particle_motion.Update_particle_position_Multiple,
            Drag_fn=lambda av, dt_: av),
        steps=STEPS)
    final = step_fn(all_variables)

    out = {"steps": np.int64(final.Step_count), "time": np.asarray(final.velocity[0].bc.time_stamp, dtype=F),
           "centers": np.asarray(final.particles.particle_center, dtype=F), "half": np.int64(HALF)}
    dx = (cfg["domain"][0][1] - cfg["domain"][0][0]) / cfg["shape"][0]
    boxes = []
    for xc, yc in b["centers"]:
        i, j = int(round(xc / dx)), int(round(yc / dx))
        boxes.append((i - HALF, i + HALF, j - HALF, j + HALF))
    out["boxes"] = np.array(boxes)
    for key, var in (("u", final.velocity[0]), ("v", final.velocity[1]), ("p", final.pressure)):
        full = np.asarray(var.data, dtype=F)
        out[key + "_crops"] = np.stack([full[a:b_, c:d] for a, b_, c, d in boxes])
        out[key + "_l2"] = np.float64(np.sqrt((full.astype(np.float64) ** 2).sum()))
        out[key + "_sum"] = np.float64(full.astype(np.float64).sum())
        # a coarse picture of the whole field: 64 x 64 block means (catches anything wrong far from the bodies)
        out[key + "_blocks"] = full.astype(np.float64).reshape(64, 32, 64, 32).mean(axis=(1, 3)).astype(F)
    path = os.path.join(HERE, "reference_ib2048_3.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t_start:.0f} s")


def main_channel():
    t_start = time.time()
    sys.dont_write_bytecode = True
    sys.path.insert(0, ROOT)
    from jax_ib_b200 import configs
    cfg = configs.channel()
    u0, v0, p0 = configs.initial_fields(cfg)
    gold = np.load(os.path.join(HERE, "channel1024x256_5.npz"))
    steps = int(gold["steps"])

    sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_standin"))
    sys.path.insert(0, REF)
    import jax.numpy as jnp
    import jax_cfd.base as cfd
    import jax_ib.base as ib
    from jax_ib.base import advection, boundaries, grids, particle_class as pc

    dt = cfg["dt"]
    grid = grids.Grid(cfg["shape"], domain=cfg["domain"])
    zero = lambda t: 0.0
    bcs = tuple(boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), tuple(w)), 0.0, [zero] * 4) for w in (cfg["wall_u"], cfg["wall_v"]))
    v = tuple(grids.GridVariable(grids.GridArray(jnp.asarray(a), off, grid), bc) for a, off, bc in zip((u0, v0), grid.cell_faces, bcs))
    p = grids.GridVariable(grids.GridArray(jnp.asarray(p0), grid.cell_center, grid), boundaries.get_pressure_bc_from_velocity(v))
    fx, fy = cfg["force"]
    forcing = lambda vv: tuple(grids.GridArray(f * jnp.ones_like(c.data), c.offset, c.grid) for f, c in zip((fx, fy), vv))
    no_bodies = lambda av, dt_: tuple(grids.GridVariable(grids.GridArray(jnp.zeros_like(c.data), c.offset, c.grid), c.bc) for c in av.velocity)
    step_fn = cfd.funcutils.repeated(
        ib.equations.semi_implicit_navier_stokes_timeBC(
            density=cfg["density"], viscosity=cfg["viscosity"], dt=dt, grid=grid,
            convect=lambda vv: tuple(advection.advect_van_leer(c, vv, dt) for c in vv),
            pressure_solve=ib.pressure.solve_fast_diag_moving_wall, forcing=forcing,
            time_stepper=ib.time_stepping.forward_euler_updated, IBM_forcing=no_bodies,
            Updating_Position=lambda av, dt_: av, Drag_fn=lambda av, dt_: av),
        steps=steps)
    final = step_fn(pc.All_Variables(None, v, p, [0], 0, [0]))
    a, b, c, d = (int(x) for x in gold["crop"])
    out = {"crop": gold["crop"], "steps": np.int64(steps), "time": np.asarray(final.velocity[0].bc.time_stamp, dtype=F)}
    for key, var in (("u", final.velocity[0]), ("v", final.velocity[1]), ("p", final.pressure)):
        full = np.asarray(var.data, dtype=F)
        out[key + "_crop"] = full[a:b, c:d]
        out[key + "_l2"] = np.float64(np.sqrt((full.astype(np.float64) ** 2).sum()))
    path = os.path.join(HERE, "reference_channel1024x256_5.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t_start:.0f} s")


def main_small64():
    """small64_50: the 64 x 64 periodic parity case with one flapping ellipse on the library's own laws, fifty steps."""
    t_start = time.time()
    sys.dont_write_bytecode = True
    sys.path.insert(0, ROOT)
    from jax_ib_b200 import configs
    cfg = configs.small_periodic(64)
    u0, v0, p0 = configs.initial_fields(cfg)
    gold = np.load(os.path.join(HERE, "small64_50.npz"))
    steps = int(gold["steps"])

    sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_standin"))
    sys.path.insert(0, REF)
    import jax.numpy as jnp
    import jax_cfd.base as cfd
    import jax_ib.base as ib
    from jax_ib.base import IBM_Force, advection, boundaries, convolution_functions, grids, kinematics as ks, particle_class as pc, particle_motion

    dt = cfg["dt"]
    grid = grids.Grid(cfg["shape"], domain=cfg["domain"])
    zero = lambda t: 0.0
    bc = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=[zero] * 4, time_stamp=0.0)
    v = tuple(grids.GridVariable(grids.GridArray(jnp.asarray(a), off, grid), bc) for a, off in zip((u0, v0), grid.cell_faces))
    p = grids.GridVariable(grids.GridArray(jnp.asarray(p0), grid.cell_center, grid), boundaries.get_pressure_bc_from_velocity(v))
    b = cfg["bodies"]
    shape_fn = lambda geometry_param, theta: tuple(jnp.asarray(a) for a in configs.ellipse_markers(geometry_param, ntheta=b["ntheta"]))
    particles = pc.particle(jnp.asarray(np.asarray(b["centers"], dtype=F)), jnp.asarray(np.asarray(b["geometry"], dtype=F)),
                            jnp.asarray(np.asarray(b["displacement_param"], dtype=F)), jnp.asarray(np.asarray(b["rotation_param"], dtype=F)),
                            pc.Grid1d(2, domain=(0, 2 * np.pi)), shape_fn, ks.displacement, ks.rotation)
    delta = lambda x, x0, w1: convolution_functions.delta_approx_logistjax(x, x0, w1)
    surf = lambda field, xp, yp: convolution_functions.new_surf_fn(field, xp, yp, delta)
    step_fn = cfd.funcutils.repeated(
        ib.equations.semi_implicit_navier_stokes_timeBC(
            density=cfg["density"], viscosity=cfg["viscosity"], dt=dt, grid=grid,
            convect=lambda vv: tuple(advection.advect_upwind(c, vv, dt) for c in vv),
            pressure_solve=ib.pressure.solve_fast_diag, forcing=None, time_stepper=ib.time_stepping.forward_euler_updated,
            IBM_forcing=lambda av, dt_: IBM_Force.calc_IBM_force_NEW_MULTIPLE(av, delta, surf, dt_),
            Updating_Position=particle_motion.Update_particle_position_Multiple, Drag_fn=lambda av, dt_: av),
        steps=steps)
    final = step_fn(pc.All_Variables(particles, v, p, [0], 0, [0]))
    a, b_, c, d = (int(x) for x in gold["crop"])
    out = {"crop": gold["crop"], "steps": np.int64(final.Step_count), "time": np.asarray(final.velocity[0].bc.time_stamp, dtype=F),
           "centers": np.asarray(final.particles.particle_center, dtype=F)}
    for key, var in (("u", final.velocity[0]), ("v", final.velocity[1]), ("p", final.pressure)):
        full = np.asarray(var.data, dtype=F)
        out[key + "_crop"] = full[a:b_, c:d]
        out[key + "_l2"] = np.float64(np.sqrt((full.astype(np.float64) ** 2).sum()))
    path = os.path.join(HERE, "reference_small64_50.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t_start:.0f} s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["channel", "small64", "ib2048"]
    if "small64" in which:
        main_small64()
    if "channel" in which:
        main_channel()
    if "ib2048" in which:
        main_ib2048()
