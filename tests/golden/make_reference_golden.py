"""Golden vectors from the REFERENCE'S OWN SOURCE FILES (run in the build container only; /root/reference is not on the
GPU box).

    python tests/golden/make_reference_golden.py            # writes tests/golden/reference_pinned.npz

hashimmg/jax_IB is pure Python on top of JAX, and JAX cannot be installed here.  oracle/jax_standin/ is a NumPy-backed
stand-in for the handful of `jax` entry points the reference's own modules use (float32 / int32 results, weakly
typed scalars, vmap / pmap as loops, jacrev as a complex-step derivative -- see its docstring).  With it on sys.path
the UNMODIFIED files /root/reference/jax_ib/base/{grids, boundaries, finite_differences, interpolation, advection,
convolution_functions, IBM_Force, kinematics, particle_class, particle_motion, array_utils, diffusion, pressure,
time_stepping, equations}.py, MD/{CFD_MD_coupling, simulate, interaction_potential}.py and penalty/{util_funs,
parameteric_fns}.py are imported -- as the reference's own package, `import jax_ib`, from /root/reference -- and executed on
seeded inputs.  Inputs and outputs
go into one .npz; tests/test_reference_pinned.py feeds the same inputs to oracle/ and compares (CPU, no reference
needed).

Two third-party dependencies of those files are absent as well and have stand-ins next to jax's:
  * tree_math (Vector / wrap / unwrap: how equations.py and time_stepping.py add pytrees) -- oracle/jax_standin/tree_math;
  * jax_cfd (google/jax-cfd; setup.py names it without a version, the reference's files are forks of its 0.2.x modules):
    only fast_diagonalization.pseudoinverse does work on this path.  oracle/jax_standin/jax_cfd restates its published
    algorithm (eigh per axis / fft of the first column, outer-summed eigenvalues, 1/lambda with a 10*eps cutoff).

What this pins: SURVEY 8a rows a1-a19 -- delta, interpolation, direct forcing + spreading, multi-body loop, kinematic
laws and their time derivatives, body update, advection schemes, diffusion, differences, ghost-cell rules, clock; the
pressure right-hand side, the 1-D operators with their boundary rows, projection, pressure accumulation and re-imposed
walls for the three boundary types; and the ASSEMBLED step (semi_implicit_navier_stokes_timeBC driven by
forward_euler_updated exactly as Flapping_Demo.ipynb builds it, one flapping body, 1 and 6 steps, three boundary types).
What it cannot pin: the eigen-solve inside jax_cfd itself (restated on both sides; anchored on identities in
tests/test_oracle_identities.py) and XLA's transcendental functions / FFT (1e-6-level differences).
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
F = np.float32


def load_reference():
    """Puts the stand-ins and the reference tree on sys.path and imports the reference's OWN package: with jax,
    tree_math, jax_cfd and jax_md standing in, `import jax_ib` runs its unmodified __init__ files."""
    sys.dont_write_bytecode = True  # /root/reference is read-only
    sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_standin"))
    sys.path.insert(0, REF)
    import jax_ib  # noqa: F401
    names = {"base": ("grids", "particle_class", "boundaries", "interpolation", "finite_differences", "advection",
                      "convolution_functions", "IBM_Force", "kinematics", "particle_motion", "array_utils", "diffusion",
                      "pressure", "time_stepping", "equations"),
             "MD": ("CFD_MD_coupling", "simulate", "interaction_potential"), "penalty": ("util_funs", "parameteric_fns")}
    ns = {n: importlib.import_module(f"jax_ib.{pkg}.{n}") for pkg, mods in names.items() for n in mods}
    assert all(m.__file__.startswith(REF + "/") for m in ns.values())
    return types.SimpleNamespace(**ns)


def ellipse_markers(a, b, n):
    """The notebooks' shape function (Flapping_Demo.ipynb cell 5) is user code, not library code: both sides get the
    same body-frame markers."""
    x = np.linspace(-a, a, n)
    y = b * np.sqrt(np.maximum(0.0, 1.0 - (x / a) ** 2))
    return np.concatenate([x, x[::-1][1:-1]]).astype(F), np.concatenate([y, -y[::-1][1:-1]]).astype(F)


def generate():
    """Runs the reference's files and returns {name: array}."""
    R = load_reference()
    import jax.numpy as jnp
    out = {}
    rng = np.random.default_rng(2024)
    shape, domain = (48, 40), ((0.5, 2.9), (0.25, 1.85))
    grid = R.grids.Grid(shape, domain=domain)
    out["shape"], out["domain"] = np.array(shape), np.array(domain)
    u, v, p = (rng.standard_normal(shape).astype(F) for _ in range(3))
    out.update(u=u, v=v, p=p)
    dt = 1.0e-3
    out["dt"] = np.float64(dt)

    def variables(bcu, bcv, bcp):
        U = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u), (1.0, 0.5), grid), bcu)
        V = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(v), (0.5, 1.0), grid), bcv)
        P = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(p), (0.5, 0.5), grid), bcp)
        return U, V, P

    # ---- a8-a13: stencils and ghost-cell rules, periodic and moving-wall channel ------------------------------
    walls = ((0.3, -0.2), (0.0, 0.0))  # u and v on the lower / upper y wall
    out["walls"] = np.array(walls)
    per = R.boundaries.periodic_boundary_conditions(2)
    bcs = {"periodic": (per, per),
           "channel": tuple(R.boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), w), 0.0, None) for w in walls)}
    far = tuple(R.boundaries.Far_field_boundary_conditions(2, w, 0.0, None) for w in (((0.5, 0.4), (0.3, -0.2)), ((0.0, 0.1), (0.0, 0.0))))
    out["far_walls"] = np.array([[[0.5, 0.4], [0.3, -0.2]], [[0.0, 0.1], [0.0, 0.0]]])
    bcs["far_field"] = far  # Dirichlet on all four walls, non-zero values on each
    for tag, (bcu, bcv) in bcs.items():
        U, V, _ = variables(bcu, bcv, per)
        P = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(p), (0.5, 0.5), grid), R.boundaries.get_pressure_bc_from_velocity((U, V)))
        vel = (U, V)
        for k, c in enumerate(vel):
            out[f"{tag}_upwind_{k}"] = np.asarray(R.advection.advect_upwind(c, vel, dt).data)
            out[f"{tag}_van_leer_{k}"] = np.asarray(R.advection.advect_van_leer(c, vel, dt).data)
            out[f"{tag}_van_leer_limiters_{k}"] = np.asarray(R.advection.advect_van_leer_using_limiters(c, vel, dt).data)
            out[f"{tag}_laplacian_{k}"] = np.asarray(R.finite_differences.laplacian(c).data)
            for axis in (0, 1):
                for s in (-2, -1, 1, 2):
                    out[f"{tag}_shift_{k}_{axis}_{s}"] = np.asarray(c.shift(s, axis).data)
        out[f"{tag}_divergence"] = np.asarray(R.finite_differences.divergence(vel).data)
        for axis in (0, 1):
            out[f"{tag}_grad_p_{axis}"] = np.asarray(R.finite_differences.forward_difference(P, axis).data)
            out[f"{tag}_back_u_{axis}"] = np.asarray(R.finite_differences.backward_difference(U, axis).data)

    # ---- a1, a2: discrete delta and Eulerian -> Lagrangian interpolation (dense, as the reference) -------------
    delta = R.convolution_functions.delta_approx_logistjax
    xs = rng.uniform(0.5, 2.9, 64).astype(F)
    out["delta_x"], out["delta_x0"], out["delta_w"] = xs, F(1.7), np.float64(grid.step[0])
    out["delta_out"] = np.asarray(delta(jnp.asarray(xs), F(1.7), grid.step[0]))
    x0, y0 = ellipse_markers(0.5, 0.06, 40)
    out["x0"], out["y0"] = x0, y0
    U, V, _ = variables(per, per, per)
    th = F(0.4)
    xp = (x0 * np.cos(th) - y0 * np.sin(th) + F(1.6)).astype(F)
    yp = (x0 * np.sin(th) + y0 * np.cos(th) + F(1.0)).astype(F)
    out["surf_xp"], out["surf_yp"] = xp, yp
    out["surf_u"] = np.asarray(R.convolution_functions.new_surf_fn(U, jnp.asarray(xp), jnp.asarray(yp), delta))
    out["surf_v"] = np.asarray(R.convolution_functions.new_surf_fn(V, jnp.asarray(xp), jnp.asarray(yp), delta))

    # ---- a6: kinematic laws and their time derivatives (jax.jacrev in the reference) ---------------------------
    K = R.kinematics
    disp_param, rot_param = [[2.8, 0.25]], [[np.pi / 2, np.pi / 4, 0.25, 0.3]]
    ts = np.array([0.0, 0.37, 1.234, 7.5], dtype=F)
    out["kin_t"] = ts
    import jax
    out["kin_disp"] = np.stack([np.asarray(K.displacement(disp_param, t)) for t in ts])
    out["kin_rot"] = np.array([np.asarray(K.rotation(rot_param, t)) for t in ts])
    out["kin_disp_dot"] = np.stack([np.asarray(jax.jacrev(lambda t: K.displacement(disp_param, t))(t)) for t in ts])
    out["kin_rot_dot"] = np.array([np.asarray(jax.jacrev(lambda t: K.rotation(rot_param, t))(t)) for t in ts])
    four = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
    out["kin_fourier_disp"] = np.stack([np.asarray(K.Displacement_Foil_Fourier_Dotted_Mutliple(four, t)) for t in ts])
    out["kin_fourier_rot"] = np.stack([np.asarray(K.rotation_Foil_Fourier_Dotted_Mutliple(four, t)) for t in ts])

    # ---- a3-a5: direct forcing + spreading for two bodies; a7: body update ------------------------------------
    t0 = F(0.37)
    pbc = R.boundaries.ConstantBoundaryConditions(values=((0.0, 0.0), (0.0, 0.0)), time_stamp=t0, types=per.types, boundary_fn=None)
    U, V, P = variables(pbc, pbc, per)
    centers = jnp.asarray(np.array([[1.2, 0.9], [2.1, 1.2]], dtype=F))
    dparams = [[0.6, 0.25], [0.4, 0.4]]
    rparams = [[np.pi / 2, np.pi / 4, 0.25, 0.0], [0.3, 0.5, 0.4, 1.0]]
    geom = [[0.5, 0.06], [0.5, 0.06]]
    shape_fn = lambda g, grid_p: (jnp.asarray(x0), jnp.asarray(y0))
    grid1d = R.particle_class.Grid1d(len(x0), domain=(0, 2 * np.pi))
    particles = R.particle_class.particle(centers, geom, dparams, rparams, grid1d, shape_fn, K.displacement, K.rotation)
    av = R.particle_class.All_Variables(particles, (U, V), P, [], 0, None)
    out["ib_centers"], out["ib_t0"] = np.asarray(centers), t0
    out["ib_dparams"], out["ib_rparams"] = np.array(dparams), np.array(rparams)
    surf = lambda field, xp_, yp_: R.convolution_functions.new_surf_fn(field, xp_, yp_, delta)
    fx, fy = R.IBM_Force.calc_IBM_force_NEW_MULTIPLE(av, delta, surf, dt)
    out["ib_fx"], out["ib_fy"] = np.asarray(fx.data), np.asarray(fy.data)
    # body update: the single-body laws take one parameter set (Flapping_Demo.ipynb), so one body here
    one = R.particle_class.particle(centers[:1], geom[:1], dparams[:1], rparams[:1], grid1d, shape_fn, K.displacement, K.rotation)
    moved = R.particle_motion.Update_particle_position_Multiple(R.particle_class.All_Variables(one, (U, V), P, [], 3, None), dt)
    out["ib_new_centers"] = np.asarray(moved.particles.particle_center)
    out["ib_step_count"] = np.int64(moved.Step_count)

    # ---- a18: the clock and the wall values that update_BC re-evaluates every step -----------------------------
    wall_u = [lambda t: 0.0, lambda t: 0.0, lambda t: 0.1 * jnp.sin(3.0 * t), lambda t: -0.2 * jnp.cos(2.0 * t)]
    wall_v = [lambda t: 0.0, lambda t: 0.0, lambda t: 0.0, lambda t: 0.0]
    bu = R.boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, -0.2)), F(0.0), wall_u)
    bv = R.boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), F(0.0), wall_v)
    U, V, P = variables(bu, bv, per)
    av = R.particle_class.All_Variables(particles, (U, V), P, [], 0, None)
    stamps, vals = [], []
    for _ in range(25):
        av = R.boundaries.update_BC(av, 5.0e-4)
        stamps.append(np.asarray(av.velocity[0].bc.time_stamp, dtype=F))
        vals.append(np.asarray([np.asarray(x, dtype=F) for x in av.velocity[0].bc.bc_values[1]] if hasattr(av.velocity[0].bc, "bc_values")
                               else [np.asarray(x, dtype=F) for x in av.velocity[0].bc._values[1]]))
    out["clock_stamps"], out["clock_wall_u"] = np.array(stamps), np.array(vals)

    # ---- a14-a17: pressure solve and projection through the reference's pressure.py / array_utils.py ----------
    # (jax_cfd's fast_diagonalization.pseudoinverse, which they call, is the stand-in's restatement)
    solvers = {"periodic": R.pressure.solve_fast_diag, "channel": R.pressure.solve_fast_diag_moving_wall,
               "far_field": R.pressure.solve_fast_diag_Far_Field}
    for tag, (bcu, bcv) in bcs.items():
        U, V, _ = variables(bcu, bcv, per)
        P = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(p), (0.5, 0.5), grid), R.boundaries.get_pressure_bc_from_velocity((U, V)))
        out[f"{tag}_q"] = np.asarray(solvers[tag]((U, V)).data)
        new = R.pressure.projection_and_update_pressure(R.particle_class.All_Variables(None, (U, V), P, [], 0, None), solvers[tag])
        out[f"{tag}_proj_u"], out[f"{tag}_proj_v"] = (np.asarray(x.data) for x in new.velocity)
        out[f"{tag}_proj_p"] = np.asarray(new.pressure.data)

    # ---- a19: the assembled step, as the notebooks build it (Flapping_Demo.ipynb cell 9) ------------------------
    # equations.semi_implicit_navier_stokes_timeBC + time_stepping.forward_euler_updated, the reference's files
    # driving the reference's stages; tree_math / jax_cfd.funcutils are the stand-in's.
    # The notebooks drive the step through jax.lax.scan (funcutils.repeated / trajectory), which turns every pytree
    # LEAF of the carried state into a float32 array -- including the clock, a leaf of the BC object
    # (boundaries.py:79-83).  The stand-in has no scan, so the clock starts as float32 here to accumulate as it does
    # there (the a18 block above pins that accumulation on its own).
    density, viscosity, nsteps = 1.3, 0.07, 6
    t_start = F(0.0)
    out["step_density"], out["step_viscosity"], out["step_n"] = np.float64(density), np.float64(viscosity), np.int64(nsteps)
    u0s, v0s = (0.3 * u).astype(F), (0.3 * v).astype(F)
    out["step_u0"], out["step_v0"] = u0s, v0s
    ib_forcing = lambda av_, dt_: R.IBM_Force.calc_IBM_force_NEW_MULTIPLE(av_, delta, surf, dt_)
    zero = lambda t: 0.0
    wall_fns = [zero, zero, lambda t: 0.1 * jnp.sin(3.0 * t), lambda t: -0.2 * jnp.cos(2.0 * t)]
    step_cases = {
        "periodic": ("upwind", R.pressure.solve_fast_diag, None,
                     (R.boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=[zero] * 4, time_stamp=t_start),) * 2),
        "channel": ("van_leer_limiters", R.pressure.solve_fast_diag_moving_wall, (0.4, -0.1),
                    (R.boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, -0.2)), t_start, wall_fns),
                     R.boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), t_start, [zero] * 4))),
        "far_field": ("upwind", R.pressure.solve_fast_diag_Far_Field, None,
                      (R.boundaries.Far_field_boundary_conditions(2, ((0.2, 0.2), (0.2, 0.2)), t_start, [lambda t: 0.2] * 4),
                       R.boundaries.Far_field_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), t_start, [zero] * 4))),
    }
    schemes = {"upwind": R.advection.advect_upwind, "van_leer_limiters": R.advection.advect_van_leer_using_limiters}
    for tag, (scheme, solver, force, (bcu, bcv)) in step_cases.items():
        U = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u0s), (1.0, 0.5), grid), bcu)
        V = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(v0s), (0.5, 1.0), grid), bcv)
        P = R.grids.GridVariable(R.grids.GridArray(jnp.zeros(shape), (0.5, 0.5), grid), R.boundaries.get_pressure_bc_from_velocity((U, V)))
        body = R.particle_class.particle(centers[:1], geom[:1], dparams[:1], rparams[:1], grid1d, shape_fn, K.displacement, K.rotation)
        av = R.particle_class.All_Variables(body, (U, V), P, [0], 0, [0])
        convect = lambda vv, fn=schemes[scheme]: tuple(fn(c, vv, dt) for c in vv)
        forcing = None if force is None else (
            lambda vv, f=force: tuple(R.grids.GridArray(fk * jnp.ones_like(c.data), c.offset, c.grid) for fk, c in zip(f, vv)))
        step = R.equations.semi_implicit_navier_stokes_timeBC(
            density=density, viscosity=viscosity, dt=dt, grid=grid, convect=convect, pressure_solve=solver,
            forcing=forcing, time_stepper=R.time_stepping.forward_euler_updated, IBM_forcing=ib_forcing,
            Updating_Position=R.particle_motion.Update_particle_position_Multiple, Drag_fn=lambda av_, dt_: av_)
        for n in range(nsteps):
            av = step(av)
            if n == 0:  # after ONE step: the order of the stages, before round-off has had time to spread
                out[f"step1_{tag}_u"], out[f"step1_{tag}_v"] = (np.asarray(x.data) for x in av.velocity)
                out[f"step1_{tag}_p"] = np.asarray(av.pressure.data)
        out[f"step_{tag}_u"], out[f"step_{tag}_v"] = (np.asarray(x.data) for x in av.velocity)
        out[f"step_{tag}_p"] = np.asarray(av.pressure.data)
        out[f"step_{tag}_centers"] = np.asarray(av.particles.particle_center)
        out[f"step_{tag}_time"] = np.asarray(av.velocity[0].bc.time_stamp, dtype=F)
        out[f"step_{tag}_count"] = np.int64(av.Step_count)

    # ---- a21: tracer coupling (MD/CFD_MD_coupling.py, MD/simulate.py brownian_cfd at kT = 0) ------------------
    import jax.random
    jax.random.ZERO_TEMPERATURE_ONLY = True
    U, V, P = variables(per, per, per)
    npts = 200
    pts = np.stack([rng.uniform(0.5, 2.9, npts), rng.uniform(0.25, 1.85, npts)], axis=1).astype(F)
    pts[:6] = np.array([[0.5, 0.25], [2.9, 1.85], [0.52, 1.0], [2.89, 0.3], [1.0, 1.849], [0.45, 0.2]], dtype=F)  # edges, outside
    out["md_pts"] = pts
    mobile = (np.arange(npts) % 5 != 0)
    out["md_mobile"] = mobile
    _, free_shift = __import__("jax_md").space.free()
    masked_shift = lambda R_, dR_, **k: jnp.where(jnp.asarray(mobile)[:, None], free_shift(R_, dR_), R_)  # the notebooks' masked shift
    init_fn, apply_fn = R.simulate.brownian_cfd(lambda R_, **k: jnp.zeros_like(R_), masked_shift, dt, 0.0,
                                               R.CFD_MD_coupling.custom_force_fn_pbc, gamma=0.8)
    md = init_fn(jax.random.PRNGKey(0), jnp.asarray(pts))
    av = R.particle_class.All_Variables(particles, (U, V), P, [], 0, md)
    out["md_force"] = np.asarray(R.CFD_MD_coupling.custom_force_fn_pbc(av))
    for _ in range(3):
        av = R.particle_class.All_Variables(particles, (U, V), P, [], 0, apply_fn(av))
    out["md_pos3"] = np.asarray(av.MD_var.position)
    dr = jnp.asarray(np.linspace(0.05, 2.0, 64).astype(F))
    out["morse_dr"] = np.asarray(dr)
    out["morse_u"] = np.asarray(R.interaction_potential.harmonic_morse(dr, h=jnp.asarray(F(5.0)), D0=jnp.asarray(F(0.7)), alpha=jnp.asarray(F(10.0)),
                                                                       r0=jnp.asarray(F(0.9)), k=jnp.asarray(F(1.0))))

    # ---- a22: penalty (Brinkman) mask and forcing (penalty/util_funs.py, parameteric_fns.py) -------------------
    theta = jnp.asarray(np.linspace(0, 2 * np.pi, 61).astype(F))
    shapes = {"ellipse": R.parameteric_fns.param_ellipse([0.45, 0.2], theta), "circle": R.parameteric_fns.param_circle([0.3], theta),
              "rose": R.parameteric_fns.param_rose([0.4, 2.0], theta)}
    smooth = lambda G_, Know: Know[0] / 2.0 * (1.0 + jnp.tanh(G_ / Know[1]))
    know = (2000.0, 0.01)
    out["pen_know"] = np.array(know)
    for name, Rth in shapes.items():
        out[f"pen_R_{name}"] = np.asarray(Rth)
        with np.errstate(all="ignore"):
            _, Rfinal = R.util_funs.project_particle(grid, (1.6, 1.0), Rth, None)
            out[f"pen_Rfinal_{name}"] = np.asarray(Rfinal)
            out[f"pen_perm_{name}"] = np.asarray(R.util_funs.calc_perm(grid, (1.6, 1.0), Rth, smooth, know))
    perm = jnp.asarray(out["pen_perm_ellipse"])
    frc = R.util_funs.arbitrary_obstacle((0.3, -0.1), perm)((U, V))
    out["pen_force_u"], out["pen_force_v"] = np.asarray(frc[0].data), np.asarray(frc[1].data)
    # the penalty stepper: semi_implicit_navier_stokes_penalty + forward_euler_penalty (equations.py:245-280,
    # time_stepping.py:259-376), channel walls, Brinkman forcing, three steps
    bu = R.boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), t_start, [zero] * 4)
    U = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u0s), (1.0, 0.5), grid), bu)
    V = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(v0s), (0.5, 1.0), grid), bu)
    P = R.grids.GridVariable(R.grids.GridArray(jnp.zeros(shape), (0.5, 0.5), grid), R.boundaries.get_pressure_bc_from_velocity((U, V)))
    pstep = R.equations.semi_implicit_navier_stokes_penalty(
        density=density, viscosity=viscosity, dt=dt, grid=grid,
        convect=lambda vv: tuple(R.advection.advect_upwind(c, vv, dt) for c in vv),
        pressure_solve=R.pressure.solve_fast_diag_moving_wall, forcing=R.util_funs.arbitrary_obstacle((0.3, -0.1), perm),
        time_stepper=R.time_stepping.forward_euler_penalty)
    av = R.particle_class.All_Variables(particles, (U, V), P, [0], 0, [0])
    for _ in range(3):
        av = pstep(av)
    out["pen_step_u"], out["pen_step_v"] = (np.asarray(x.data) for x in av.velocity)
    out["pen_step_p"] = np.asarray(av.pressure.data)
    out["pen_step_time"] = np.asarray(av.velocity[0].bc.time_stamp, dtype=F)

    # ---- a4-a7 with the multi-body Fourier laws of the journal-bearing notebook (two bodies) --------------------
    four_d = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
    four_r = [[0.8, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.3, 1.0], [0.5, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], -0.7, 0.5]]
    out["kin_fourier_norm_rot"] = np.stack([np.asarray(K.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED(four_r, t)) for t in ts])
    out["kin_fourier_disp_dot"] = np.stack([np.asarray(jax.jacrev(lambda t: K.Displacement_Foil_Fourier_Dotted_Mutliple(four_d, t))(t)) for t in ts])
    out["kin_fourier_norm_rot_dot"] = np.stack([np.asarray(jax.jacrev(lambda t: K.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED(four_r, t))(t)) for t in ts])
    U = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u0s), (1.0, 0.5), grid), R.boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=[zero] * 4, time_stamp=F(0.37)))
    V = R.grids.GridVariable(R.grids.GridArray(jnp.asarray(v0s), (0.5, 1.0), grid), U.bc)
    P = R.grids.GridVariable(R.grids.GridArray(jnp.zeros(shape), (0.5, 0.5), grid), R.boundaries.get_pressure_bc_from_velocity((U, V)))
    two = R.particle_class.particle(centers, geom, four_d, four_r, grid1d, shape_fn, K.Displacement_Foil_Fourier_Dotted_Mutliple,
                                    K.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED)
    av = R.particle_class.All_Variables(two, (U, V), P, [0], 0, [0])
    fx, fy = R.IBM_Force.calc_IBM_force_NEW_MULTIPLE(av, delta, surf, dt)
    out["two_fx"], out["two_fy"] = np.asarray(fx.data), np.asarray(fy.data)
    step = R.equations.semi_implicit_navier_stokes_timeBC(
        density=density, viscosity=viscosity, dt=dt, grid=grid, convect=lambda vv: tuple(R.advection.advect_upwind(c, vv, dt) for c in vv),
        pressure_solve=R.pressure.solve_fast_diag, forcing=None, time_stepper=R.time_stepping.forward_euler_updated,
        IBM_forcing=ib_forcing, Updating_Position=R.particle_motion.Update_particle_position_Multiple, Drag_fn=lambda av_, dt_: av_)
    for _ in range(3):
        av = step(av)
    out["two_step_u"], out["two_step_v"] = (np.asarray(x.data) for x in av.velocity)
    out["two_step_p"] = np.asarray(av.pressure.data)
    out["two_step_centers"] = np.asarray(av.particles.particle_center)

    # ---- post-processing integral (IBM_Force.py:5-18) ----------------------------------------------------------
    U, V, P = variables(per, per, per)
    out["trapz_u"] = np.asarray(R.IBM_Force.Integrate_Field_Fluid_Domain(U))

    return out


def main():
    out = generate()
    path = os.path.join(HERE, "reference_pinned.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
