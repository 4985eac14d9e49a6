"""The CUDA path against outputs of the REFERENCE'S OWN SOURCE FILES -- no oracle in between.

tests/golden/reference_pinned.npz holds inputs and outputs of hashimmg/jax_IB's unmodified modules, executed in the build
container on the NumPy stand-in for jax (tests/golden/make_reference_golden.py).  Here the same inputs go through the
reference-shaped API of this package (`jax_ib_b200` in place of `jax_ib.base`, wired exactly as the generator and the
notebooks wire the reference) and the results are compared with what the reference's files returned.

Bars: the north star's rel-L2 <= 1e-5 for velocities after N steps; the stored pressure of these short runs is ~50x
smaller than the velocity and carries the same absolute round-off, so its bar is 1e-4.  The reference sums every marker
over the WHOLE grid; the kernels truncate the Gaussian at `ib_window` cells (6e-9 of the sum), which is inside the bars.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import jax_ib_b200 as ib  # noqa: E402
from jax_ib_b200 import (IBM_Force, MD, advection, boundaries, convolution_functions, equations, finite_differences, funcutils,  # noqa: E402
                         grids, kinematics as ks, particle_class as pc, particle_motion, pressure)

from _adapter import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu
F = np.float32
DEV = "cuda"
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pinned.npz"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def _grid():
    return grids.Grid(tuple(int(x) for x in GOLD["shape"]), domain=tuple(tuple(float(y) for y in x) for x in GOLD["domain"]))


ZERO = [lambda t: 0.0] * 4


def _bcs(tag, t0=0.0):
    """The boundary objects of generator parts 1 and 2 (constant walls)."""
    if tag == "periodic":
        bc = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=ZERO, time_stamp=F(t0))
        return bc, bc
    if tag == "channel":
        (ulo, uhi), (vlo, vhi) = (tuple(float(x) for x in r) for r in GOLD["walls"])
        return (boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (ulo, uhi)), F(t0), [ZERO[0], ZERO[0], lambda t: ulo, lambda t: uhi]),
                boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (vlo, vhi)), F(t0), [ZERO[0], ZERO[0], lambda t: vlo, lambda t: vhi]))
    wu, wv = (tuple(tuple(float(x) for x in r) for r in w) for w in GOLD["far_walls"])
    mk = lambda w: boundaries.Far_field_boundary_conditions(2, w, F(t0), [lambda t, a=w[0][0]: a, lambda t, a=w[0][1]: a,
                                                                         lambda t, a=w[1][0]: a, lambda t, a=w[1][1]: a])
    return mk(wu), mk(wv)


def _variables(u, v, p, bcu, bcv):
    g = _grid()
    U = grids.GridVariable(grids.GridArray(dev(u), (1.0, 0.5), g), bcu)
    V = grids.GridVariable(grids.GridArray(dev(v), (0.5, 1.0), g), bcv)
    P = grids.GridVariable(grids.GridArray(dev(p), (0.5, 0.5), g), boundaries.get_pressure_bc_from_velocity((U, V)))
    return U, V, P


def _body(n, disp=ks.displacement, rot=ks.rotation, dparams=None, rparams=None):
    x0, y0 = GOLD["x0"], GOLD["y0"]
    shape_fn = lambda geometry_param, theta: (x0, y0)
    dparams = GOLD["ib_dparams"][:n].astype(F) if dparams is None else dparams
    rparams = GOLD["ib_rparams"][:n].astype(F) if rparams is None else rparams
    return pc.particle(GOLD["ib_centers"][:n].astype(F), np.array([[0.5, 0.06]] * n, F), dparams, rparams,
                       pc.Grid1d(len(x0), domain=(0, 2 * np.pi)), shape_fn, disp, rot)


DELTA = lambda x, x0, w: convolution_functions.delta_approx_logistjax(x, x0, w)
SURF = lambda field, xp, yp: convolution_functions.new_surf_fn(field, xp, yp, DELTA)
IBM = lambda av, dt: IBM_Force.calc_IBM_force_NEW_MULTIPLE(av, DELTA, SURF, dt)


# ------------------------------------------------------------------------------------------------ stages
@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_advection_and_differences_against_the_reference_files(tag):
    """advection.py:120-333 and finite_differences.py on the reference's arrays (random fields, moving-wall values)."""
    bcu, bcv = _bcs(tag)
    U, V, P = _variables(GOLD["u"], GOLD["v"], GOLD["p"], bcu, bcv)
    vel, dt = (U, V), float(GOLD["dt"])
    for k, c in enumerate(vel):
        for name, fn in (("upwind", advection.advect_upwind), ("van_leer", advection.advect_van_leer),
                         ("van_leer_limiters", advection.advect_van_leer_using_limiters)):
            want = GOLD[f"{tag}_{name}_{k}"]
            got = host(fn(c, vel, dt).data)
            assert np.abs(got - want).max() < 2e-5 * np.abs(want).max(), (tag, name, k)
        want = GOLD[f"{tag}_laplacian_{k}"]
        assert np.abs(host(finite_differences.laplacian(c).data) - want).max() < 1e-5 * np.abs(want).max()
    want = GOLD[f"{tag}_divergence"]
    assert np.abs(host(finite_differences.divergence(vel).data) - want).max() < 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_pressure_projection_against_the_reference_files(tag):
    """pressure.py:58-154 run from the reference's file: q, v - grad q with the walls re-imposed, p + q."""
    bcu, bcv = _bcs(tag)
    U, V, P = _variables(GOLD["u"], GOLD["v"], GOLD["p"], bcu, bcv)
    solve = {"periodic": pressure.solve_fast_diag, "channel": pressure.solve_fast_diag_moving_wall,
             "far_field": pressure.solve_fast_diag_Far_Field}[tag]
    # white-noise input: the float32 round-off floor of any 48 x 40 solve is ~1e-6 of q; two different transforms
    assert rel_l2(host(solve((U, V)).data), GOLD[f"{tag}_q"]) < 2e-5
    new = pressure.projection_and_update_pressure(pc.All_Variables(None, (U, V), P, [], 0, None), solve)
    assert rel_l2(host(new.velocity[0].data), GOLD[f"{tag}_proj_u"]) < 5e-6
    assert rel_l2(host(new.velocity[1].data), GOLD[f"{tag}_proj_v"]) < 5e-6
    assert rel_l2(host(new.pressure.data), GOLD[f"{tag}_proj_p"]) < 5e-6
    if tag != "periodic":  # the wall-aligned column of v holds the wall value exactly (impose_bc, pressure.py:87-89)
        np.testing.assert_array_equal(host(new.velocity[1].data)[:, -1], GOLD[f"{tag}_proj_v"][:, -1])
    if tag == "far_field":
        np.testing.assert_array_equal(host(new.velocity[0].data)[-1, :], GOLD[f"{tag}_proj_u"][-1, :])


def test_direct_forcing_of_two_bodies_against_the_reference_files():
    """IBM_Force.py:20-115 with the reference's DENSE sums as the target; the grid starts at (0.5, 0.25)."""
    bcu, bcv = _bcs("periodic", float(GOLD["ib_t0"]))
    U, V, P = _variables(GOLD["u"], GOLD["v"], GOLD["p"], bcu, bcv)
    av = pc.All_Variables(_body(2), (U, V), P, [], 0, None)
    fx, fy = IBM(av, float(GOLD["dt"]))
    assert rel_l2(host(fx.data), GOLD["ib_fx"]) < 5e-6
    assert rel_l2(host(fy.data), GOLD["ib_fy"]) < 5e-6
    got = SURF(U, GOLD["surf_xp"], GOLD["surf_yp"])
    np.testing.assert_allclose(host(got), GOLD["surf_u"], rtol=5e-6, atol=2e-6)


def test_per_component_forcing_entry_points_against_the_reference_files():
    """IBM_Force.py:89-106 IBM_Multiple_NEW (all bodies, one component) on top of IBM_force_GENERAL (:20-87, one body,
    caller-supplied kinematic callables), and the domain integral :5-18."""
    bcu, bcv = _bcs("periodic", float(GOLD["ib_t0"]))
    U, V, P = _variables(GOLD["u"], GOLD["v"], GOLD["p"], bcu, bcv)
    two = _body(2)
    fx = IBM_Force.IBM_Multiple_NEW(U, 0, two, DELTA, SURF, float(GOLD["dt"]))
    fy = IBM_Force.IBM_Multiple_NEW(V, 1, two, DELTA, SURF, float(GOLD["dt"]))
    assert tuple(fx.offset) == (1.0, 0.5) and tuple(fy.offset) == (0.5, 1.0)
    assert rel_l2(host(fx.data), GOLD["ib_fx"]) < 5e-6
    assert rel_l2(host(fy.data), GOLD["ib_fy"]) < 5e-6
    assert float(IBM_Force.Integrate_Field_Fluid_Domain(U)) == pytest.approx(float(GOLD["trapz_u"]), rel=2e-5, abs=1e-6)


def test_tracer_coupling_against_the_reference_files():
    """MD/CFD_MD_coupling.py:7-27 (points on the edges and outside the grid included) and three zero-temperature
    Brownian steps of MD/simulate.py:81-97 with the notebooks' masked shift, in a frozen velocity field."""
    import dataclasses
    bcu, bcv = _bcs("periodic")
    U, V, P = _variables(GOLD["u"], GOLD["v"], GOLD["p"], bcu, bcv)
    av = pc.All_Variables(None, (U, V), P, [], 0, MD.simulate.BrownianState(dev(GOLD["md_pts"]), 1.0, None))
    got = host(MD.CFD_MD_coupling.custom_force_fn_pbc(av))
    np.testing.assert_allclose(got, GOLD["md_force"], rtol=0, atol=2e-6)
    mobile = torch.as_tensor(GOLD["md_mobile"], device=DEV)
    masked_shift = lambda R, dR, **kw: torch.where(mobile[:, None], R + dR, R)
    _, apply_fn = MD.simulate.brownian_cfd(lambda R, **kw: torch.zeros_like(R), masked_shift, float(GOLD["dt"]), 0.0,
                                           MD.CFD_MD_coupling.custom_force_fn_pbc, gamma=0.8)
    for _ in range(3):
        av = dataclasses.replace(av, MD_var=apply_fn(av))
    np.testing.assert_allclose(host(av.MD_var.position), GOLD["md_pos3"], rtol=0, atol=5e-7)


# ------------------------------------------------------------------------------------------------ the assembled step
def _step_case(tag):
    g = _grid()
    dt, density, viscosity = float(GOLD["dt"]), float(GOLD["step_density"]), float(GOLD["step_viscosity"])
    t0 = F(0.0)
    if tag == "periodic":
        bcu = bcv = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=ZERO, time_stamp=t0)
        adv, solve, forcing = advection.advect_upwind, pressure.solve_fast_diag, None
    elif tag == "channel":
        wall = [ZERO[0], ZERO[0], lambda t: 0.1 * np.sin(F(3.0) * F(t)), lambda t: -0.2 * np.cos(F(2.0) * F(t))]
        bcu = boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, -0.2)), t0, wall)
        bcv = boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), t0, ZERO)
        adv, solve = advection.advect_van_leer_using_limiters, pressure.solve_fast_diag_moving_wall
        forcing = equations.constant_forcing((0.4, -0.1))
    else:
        bcu = boundaries.Far_field_boundary_conditions(2, ((0.2, 0.2), (0.2, 0.2)), t0, [lambda t: 0.2] * 4)
        bcv = boundaries.Far_field_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), t0, ZERO)
        adv, solve, forcing = advection.advect_upwind, pressure.solve_fast_diag_Far_Field, None
    U, V, P = _variables(GOLD["step_u0"], GOLD["step_v0"], np.zeros(g.shape, F), bcu, bcv)
    step_fn = equations.semi_implicit_navier_stokes_timeBC(
        density=density, viscosity=viscosity, dt=dt, grid=g, convect=lambda v: tuple(adv(c, v, dt) for c in v),
        pressure_solve=solve, forcing=forcing, time_stepper=ib.time_stepping.forward_euler_updated, IBM_forcing=IBM,
        Updating_Position=particle_motion.Update_particle_position_Multiple, Drag_fn=lambda av, dt_: av)
    return step_fn, pc.All_Variables(_body(1), (U, V), P, [0], 0, [0])


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_assembled_step_against_the_reference_files(tag):
    """equations.py:194-242 + time_stepping.py:109-256 as Flapping_Demo.ipynb cell 9 wires them, one flapping body:
    the reference's files produced step1_* (one step) and step_* (six steps) from the same initial state."""
    step_fn, av = _step_case(tag)
    assert step_fn.fused
    one = step_fn(av)
    assert rel_l2(host(one.velocity[0].data), GOLD[f"step1_{tag}_u"]) < 5e-6
    assert rel_l2(host(one.velocity[1].data), GOLD[f"step1_{tag}_v"]) < 5e-6
    assert rel_l2(host(one.pressure.data), GOLD[f"step1_{tag}_p"]) < 1e-4
    n = int(GOLD["step_n"])
    out = funcutils.repeated(step_fn, n)(av)
    assert rel_l2(host(out.velocity[0].data), GOLD[f"step_{tag}_u"]) < 1e-5
    assert rel_l2(host(out.velocity[1].data), GOLD[f"step_{tag}_v"]) < 1e-5
    assert rel_l2(host(out.pressure.data), GOLD[f"step_{tag}_p"]) < 1e-4
    np.testing.assert_allclose(np.asarray(out.particles.particle_center), GOLD[f"step_{tag}_centers"], rtol=0, atol=5e-7)
    assert F(out.velocity[0].bc.time_stamp) == GOLD[f"step_{tag}_time"]
    assert out.Step_count == int(GOLD[f"step_{tag}_count"])


def test_two_bodies_on_the_journal_bearing_laws_against_the_reference_files():
    """Two bodies, Fourier displacement and NORMALISED Fourier rotation (kinematics.py:14-105), three steps."""
    four_d = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
    four_r = [[0.8, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.3, 1.0], [0.5, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], -0.7, 0.5]]
    g = _grid()
    dt = float(GOLD["dt"])
    bc = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=ZERO, time_stamp=F(0.37))
    U, V, P = _variables(GOLD["step_u0"], GOLD["step_v0"], np.zeros(g.shape, F), bc, bc)
    two = _body(2, ks.Displacement_Foil_Fourier_Dotted_Mutliple, ks.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED, four_d, four_r)
    av = pc.All_Variables(two, (U, V), P, [0], 0, [0])
    fx, fy = IBM(av, dt)
    assert rel_l2(host(fx.data), GOLD["two_fx"]) < 5e-6 and rel_l2(host(fy.data), GOLD["two_fy"]) < 5e-6
    step_fn = equations.semi_implicit_navier_stokes_timeBC(
        density=float(GOLD["step_density"]), viscosity=float(GOLD["step_viscosity"]), dt=dt, grid=g,
        convect=lambda v: tuple(advection.advect_upwind(c, v, dt) for c in v), pressure_solve=pressure.solve_fast_diag,
        forcing=None, time_stepper=ib.time_stepping.forward_euler_updated, IBM_forcing=IBM,
        Updating_Position=particle_motion.Update_particle_position_Multiple, Drag_fn=lambda a, dt_: a)
    out = funcutils.repeated(step_fn, 3)(av)
    assert rel_l2(host(out.velocity[0].data), GOLD["two_step_u"]) < 1e-5
    assert rel_l2(host(out.velocity[1].data), GOLD["two_step_v"]) < 1e-5
    assert rel_l2(host(out.pressure.data), GOLD["two_step_p"]) < 1e-4
    np.testing.assert_allclose(np.asarray(out.particles.particle_center), GOLD["two_step_centers"], rtol=0, atol=1e-6)
