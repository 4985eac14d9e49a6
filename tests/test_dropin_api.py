"""The reference-facing API: the same wiring as the reference's notebooks (Flapping_Demo.ipynb cells
3/5/8, journal_bearing_demo.ipynb cells 4-10), with `jax_ib_b200` in place of `jax_ib.base`."""
import math

import numpy as np
import pytest
import torch

import jax_ib_b200 as ib
from jax_ib_b200 import (IBM_Force, MD, advection, boundaries, configs, convolution_functions, equations, funcutils, grids,
                         kinematics as ks, particle_class as pc, particle_motion)

from _adapter import oracle_state, oracle_step, rel_l2

F = np.float32


def flapping_setup(size=(600, 600), domain=((0, 15.0), (0, 15.0)), dt=5e-4, device="cpu"):
    """Flapping_Demo.ipynb cells 3, 5 and 8, line for line (jnp -> torch / numpy)."""
    density, viscosity = 1.0, 0.05
    grid = grids.Grid(size, domain=domain)

    def Boundary_fn(t):
        return 0.0

    bc_fns = [Boundary_fn] * 4
    vx_bc = ((bc_fns[0](0.0), bc_fns[1](0.0)), (bc_fns[2](0.0), bc_fns[3](0.0)))
    vy_bc = ((0.0, 0.0), (0, 0.0))
    velocity_bc = (boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=vx_bc, bc_fn=bc_fns, time_stamp=0.0),
                   boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=vy_bc, bc_fn=bc_fns, time_stamp=0.0))

    def convect(v):
        return tuple(advection.advect_upwind(u, v, dt) for u in v)

    vx_fn = lambda x, y: torch.zeros_like(x + y)
    vy_fn = lambda x, y: torch.zeros_like(x + y)
    v0 = tuple(grid.eval_on_mesh(fn, offset) for fn, offset in zip((vx_fn, vy_fn), grid.cell_faces))
    v0 = tuple(grids.GridVariable(u.to(device), bc) for u, bc in zip(v0, velocity_bc))
    pressure0 = grids.GridVariable(grids.GridArray(torch.zeros_like(v0[0].data), grid.cell_center, grid),
                                   boundaries.get_pressure_bc_from_velocity(v0))

    def foil_XY_ELLIPSE(geometry_param, theta):
        return configs.ellipse_markers(geometry_param, ntheta=150)

    particles = pc.particle(np.array([[domain[0][1] * 0.75, domain[1][1] / 2]], F), np.array([[0.5, 0.06]], F),
                            np.array([[2.8, 0.25]], F), np.array([[math.pi / 2, math.pi / 4, 0.25, 0]], F),
                            pc.Grid1d(2, domain=(0, 2 * math.pi)), foil_XY_ELLIPSE, ks.displacement, ks.rotation)
    all_variables = pc.All_Variables(particles, v0, pressure0, [0], 0, [0])

    def internal_post_processing(all_variables, dt):
        return all_variables

    discrete_delta = lambda x, x0, w1: convolution_functions.delta_approx_logistjax(x, x0, w1)
    surf_fn = lambda field, xp, yp: convolution_functions.new_surf_fn(field, xp, yp, discrete_delta)
    IBM_forcing = lambda v, dt: IBM_Force.calc_IBM_force_NEW_MULTIPLE(v, discrete_delta, surf_fn, dt)
    step_fn = equations.semi_implicit_navier_stokes_timeBC(
        density=density, viscosity=viscosity, dt=dt, grid=grid, convect=convect,
        pressure_solve=ib.pressure.solve_fast_diag, forcing=None, time_stepper=ib.time_stepping.forward_euler_updated,
        IBM_forcing=IBM_forcing, Updating_Position=particle_motion.Update_particle_position_Multiple,
        Drag_fn=internal_post_processing)
    return step_fn, all_variables


def test_notebook_wiring_is_recognised_as_the_fused_path():
    step_fn, _ = flapping_setup()
    assert step_fn.fused
    assert step_fn.cfg["convect"] == "upwind" and step_fn.cfg["poisson"] == "periodic"
    assert step_fn.cfg["ib"] and step_fn.cfg["move"] and step_fn.cfg["window"] == 6


def test_opaque_plugins_fall_back_to_the_staged_path():
    grid = grids.Grid((32, 32), domain=((0, 1.0), (0, 1.0)))
    custom = lambda v: tuple(advection.advect_upwind(u, v, 1e-3) * 0.5 for u in v)  # user arithmetic on top
    step_fn = equations.semi_implicit_navier_stokes_timeBC(1.0, 0.1, 1e-3, grid, convect=custom)
    assert not step_fn.fused
    default = equations.semi_implicit_navier_stokes_timeBC(1.0, 0.1, 1e-3, grid)
    assert default.fused and default.cfg["convect"] == "van_leer_limiters"  # equations.py:55-58
    pen = equations.semi_implicit_navier_stokes_penalty(1.0, 0.1, 1e-3, grid, forcing=equations.constant_forcing((1e-3, 0)))
    assert pen.fused and not pen.cfg["incremental"] and pen.cfg["force"] == (1e-3, 0.0)
    with pytest.raises(NotImplementedError):
        IBM_Force.calc_IBM_force_NEW_MULTIPLE(None, lambda x, x0, w: x, lambda f, a, b: a, 1e-3)


@pytest.mark.gpu
def test_flapping_demo_through_the_reference_api():
    step_fn, all_variables = flapping_setup()
    inner = funcutils.repeated(step_fn, steps=10)
    rollout = funcutils.trajectory(inner, 2, start_with_input=True)
    final, traj = rollout(all_variables)
    assert len(traj) == 2 and traj[0].Step_count == 0 and traj[1].Step_count == 10 and final.Step_count == 20
    gold = np.load(__file__.replace("test_dropin_api.py", "golden/flap600_20.npz"))
    a, b, c, d = (int(x) for x in gold["crop"])
    for key, var in (("u", final.velocity[0]), ("v", final.velocity[1]), ("p", final.pressure)):
        bar = max(1e-5, 2.0 * float(gold[key + "_noise"]))
        assert rel_l2(var.data.cpu().numpy()[a:b, c:d], gold[key + "_crop"]) < bar, key
    np.testing.assert_allclose(np.asarray(final.particles.particle_center), gold["centers"], atol=2e-6)
    assert np.float32(final.velocity[0].bc.time_stamp) == gold["time"]


@pytest.mark.gpu
def test_staged_path_matches_fused_path():
    """The same physics through the plug-in-by-plug-in path (opaque Drag_fn) and the fused path."""
    cfg = configs.small_periodic(64)
    av = ib.demo.make_all_variables(cfg, device="cuda")
    dt, grid = cfg["dt"], av.velocity[0].grid
    convect = lambda v: tuple(advection.advect_upwind(u, v, dt) for u in v)
    dd = lambda x, x0, w: convolution_functions.delta_approx_logistjax(x, x0, w)
    sf = lambda f, xp, yp: convolution_functions.new_surf_fn(f, xp, yp, dd)
    ibm = lambda v, dt: IBM_Force.calc_IBM_force_NEW_MULTIPLE(v, dd, sf, dt)
    calls = []

    def drag(all_variables, dt):  # a genuine in-loop reduction: forces the staged path
        calls.append(float(all_variables.velocity[0].data.abs().max()))
        return all_variables

    kw = dict(density=cfg["density"], viscosity=cfg["viscosity"], dt=dt, grid=grid, convect=convect,
              pressure_solve=ib.pressure.solve_fast_diag, IBM_forcing=ibm,
              Updating_Position=particle_motion.Update_particle_position_Multiple)
    fused = equations.semi_implicit_navier_stokes_timeBC(**kw)
    staged = equations.semi_implicit_navier_stokes_timeBC(Drag_fn=drag, **kw)
    assert fused.fused and not staged.fused
    a = funcutils.repeated(fused, 3)(av)
    b = funcutils.repeated(staged, 3)(av)
    assert len(calls) == 3
    for x, y in ((a.velocity[0], b.velocity[0]), (a.velocity[1], b.velocity[1]), (a.pressure, b.pressure)):
        assert rel_l2(x.data.cpu().numpy(), y.data.cpu().numpy()) < 5e-6
    np.testing.assert_allclose(np.asarray(a.particles.particle_center), np.asarray(b.particles.particle_center), atol=1e-6)
    assert a.Step_count == b.Step_count == 3


@pytest.mark.gpu
def test_journal_bearing_with_tracers_kT0():
    """journal_bearing_demo.ipynb cells 4-10 at the notebook's tracer count (1200 mobile + 200 wall + 1 hub =
    1401, :214-219), kT = 0, against the oracle."""
    from oracle import tracers

    cfg = configs.bearing300()
    av = ib.demo.make_all_variables(cfg, device="cuda")
    dt, grid = cfg["dt"], av.velocity[0].grid
    rng = np.random.default_rng(0)
    n_mob, n_wall = 1200, 200
    th = rng.uniform(0, 2 * np.pi, n_mob)
    rr = rng.uniform(0.6, 0.9, n_mob)
    mob = np.stack([2.5 + rr * np.cos(th), 2.5 + rr * np.sin(th)], 1)
    tw = np.linspace(0, 2 * np.pi, n_wall)
    wall = np.stack([2.5 + np.cos(tw), 2.5 + np.sin(tw)], 1)
    Pos_all = np.concatenate([mob, wall, [[2.5, 2.5]]]).astype(F)
    species = np.concatenate([np.zeros(n_mob), np.ones(n_wall), [2]]).astype(np.int32)
    is_mobile = np.concatenate([np.ones(n_mob), np.zeros(n_wall + 1)]).astype(bool)
    Rad_outer = 2 * np.pi / n_wall
    r0_array = np.array([[0.0, Rad_outer, 0.5], [Rad_outer, 0, 0], [0.5, 0, 0]], F)
    h_array = np.array([[0.0, 5, 5], [5, 0, 0], [5, 0, 0]], F)
    displacement, shift = MD.space.periodic(5.0, wrapped=False)
    energy_fn = MD.interaction_potential.harmonic_morse_pair(displacement, species=species, h=h_array, D0=0.0, alpha=10.0,
                                                            r0=r0_array, k=1.0)
    force_fn = MD.quantity.force(energy_fn)
    mobile_t = torch.as_tensor(is_mobile, device="cuda")

    def masked_shift(R, dR, is_mobile=None):
        return torch.where(mobile_t[:, None], shift(R, dR), R)

    init_fn, apply_fn = MD.simulate.brownian_cfd(lambda R, **kw: force_fn(R), masked_shift, dt, 0.0,
                                                 MD.CFD_MD_coupling.custom_force_fn_pbc, gamma=1.0)
    md_state = init_fn(2, torch.from_numpy(Pos_all), mass=1.0)
    step_fn_MD = lambda state: apply_fn(state, is_mobile=is_mobile)
    Update_position = lambda v, dt: particle_motion.Update_particle_position_Multiple_and_MD_Step(step_fn_MD, v, dt)
    convect = lambda v: tuple(advection.advect_upwind(u, v, dt) for u in v)
    dd = lambda x, x0, w: convolution_functions.delta_approx_logistjax(x, x0, w)
    sf = lambda f, xp, yp: convolution_functions.new_surf_fn(f, xp, yp, dd)
    step_fn = equations.semi_implicit_navier_stokes_timeBC(
        density=1.0, viscosity=0.05, dt=dt, grid=grid, convect=convect, pressure_solve=ib.pressure.solve_fast_diag,
        IBM_forcing=lambda v, dt: IBM_Force.calc_IBM_force_NEW_MULTIPLE(v, dd, sf, dt), Updating_Position=Update_position,
        Drag_fn=lambda a, dt: a)
    assert step_fn.fused and step_fn.cfg["md_step"] is step_fn_MD
    import dataclasses

    av = dataclasses.replace(av, MD_var=md_state)
    out = funcutils.repeated(step_fn, 5)(av)
    # oracle with the tracer sub-step
    st = oracle_state(cfg)
    st.MD_var = tracers.BrownianState(Pos_all.copy())
    pf = lambda R: tracers.harmonic_morse_force(R.astype(np.float64), species, h_array, r0_array, (5.0, 5.0)).astype(F)
    md = lambda s: tracers.brownian_step(s.MD_var, s.velocity, dt, 1.0, 0.0, pf, is_mobile)
    ostep = oracle_step(cfg, ib_window=6, md_step=md)
    for _ in range(5):
        st = ostep(st)
    np.testing.assert_allclose(out.MD_var.position.cpu().numpy(), st.MD_var.position, rtol=0, atol=2e-6)
    assert rel_l2(out.velocity[0].data.cpu().numpy(), st.velocity[0].data) < 1e-5
