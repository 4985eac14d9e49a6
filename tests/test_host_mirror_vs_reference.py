"""The HOST side of the drop-in (the torch mirror of the reference's data model: grids / boundaries / finite_differences /
kinematics / particle_motion / update_BC) against outputs of the reference's own files (tests/golden/reference_pinned.npz).
CPU only: these are the members of All_Variables and the plug-in fallbacks that never touch a kernel."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from jax_ib_b200 import boundaries, finite_differences as fd, grids, kinematics as ks, particle_class as pc, particle_motion  # noqa: E402

F = np.float32
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pinned.npz"))
ZERO = [lambda t: 0.0] * 4


def _grid():
    return grids.Grid(tuple(int(x) for x in GOLD["shape"]), domain=tuple(tuple(float(y) for y in x) for x in GOLD["domain"]))


def _bcs(tag):
    if tag == "periodic":
        bc = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=ZERO, time_stamp=F(0.0))
        return bc, bc
    if tag == "channel":
        (ulo, uhi), (vlo, vhi) = (tuple(float(x) for x in r) for r in GOLD["walls"])
        return (boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (ulo, uhi)), F(0.0), None),
                boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (vlo, vhi)), F(0.0), None))
    wu, wv = (tuple(tuple(float(x) for x in r) for r in w) for w in GOLD["far_walls"])
    return boundaries.Far_field_boundary_conditions(2, wu, F(0.0), None), boundaries.Far_field_boundary_conditions(2, wv, F(0.0), None)


def _variables(tag):
    g = _grid()
    bcu, bcv = _bcs(tag)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    U = grids.GridVariable(grids.GridArray(t(GOLD["u"]), (1.0, 0.5), g), bcu)
    V = grids.GridVariable(grids.GridArray(t(GOLD["v"]), (0.5, 1.0), g), bcv)
    P = grids.GridVariable(grids.GridArray(t(GOLD["p"]), (0.5, 0.5), g), boundaries.get_pressure_bc_from_velocity((U, V)))
    return U, V, P


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_ghost_cells_and_differences_equal_the_reference_bit_for_bit(tag):
    """grids.py shift, boundaries.py:95-301 pad / trim, finite_differences.py:87-146."""
    U, V, P = _variables(tag)
    for k, c in enumerate((U, V)):
        for axis in (0, 1):
            for s in (-2, -1, 1, 2):
                np.testing.assert_array_equal(c.shift(s, axis).data.numpy(), GOLD[f"{tag}_shift_{k}_{axis}_{s}"])
        np.testing.assert_array_equal(fd.laplacian(c).data.numpy(), GOLD[f"{tag}_laplacian_{k}"])
    np.testing.assert_array_equal(fd.divergence((U, V)).data.numpy(), GOLD[f"{tag}_divergence"])
    for axis in (0, 1):
        np.testing.assert_array_equal(fd.forward_difference(P, axis).data.numpy(), GOLD[f"{tag}_grad_p_{axis}"])
        np.testing.assert_array_equal(fd.backward_difference(U, axis).data.numpy(), GOLD[f"{tag}_back_u_{axis}"])


def test_kinematic_laws_match_the_reference():
    """kinematics.py:5-105 as the host evaluates them for the per-step body table."""
    disp_param, rot_param = [[2.8, 0.25]], [[np.pi / 2, np.pi / 4, 0.25, 0.3]]
    four = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
    four_r = [[0.8, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.3, 1.0], [0.5, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], -0.7, 0.5]]
    for i, t in enumerate(GOLD["kin_t"]):
        np.testing.assert_allclose(np.asarray(ks.displacement(disp_param, t)), GOLD["kin_disp"][i], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(np.asarray(ks.rotation(rot_param, t)), GOLD["kin_rot"][i], rtol=1e-6)
        np.testing.assert_allclose(np.asarray(ks.Displacement_Foil_Fourier_Dotted_Mutliple(four, t)), GOLD["kin_fourier_disp"][i], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(np.asarray(ks.rotation_Foil_Fourier_Dotted_Mutliple(four, t)), GOLD["kin_fourier_rot"][i], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(np.asarray(ks.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED(four_r, t)), GOLD["kin_fourier_norm_rot"][i],
                                   rtol=3e-6, atol=1e-6)


def test_body_table_derivatives_match_jacrev_in_the_reference():
    """IBM_Force.py:101-102 / particle_motion.py:47-52 differentiate the laws with jax.jacrev; the host builds the
    per-step table [cos, sin, xc, yc, Ux, Uy, omega, -] from closed forms or its own derivative."""
    x0, y0 = GOLD["x0"], GOLD["y0"]
    four = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
    four_r = [[0.8, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.3, 1.0], [0.5, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], -0.7, 0.5]]
    two = pc.particle(GOLD["ib_centers"].astype(F), np.array([[0.5, 0.06]] * 2, F), four, four_r, pc.Grid1d(len(x0), domain=(0, 2 * np.pi)),
                      lambda gp, th: (x0, y0), ks.Displacement_Foil_Fourier_Dotted_Mutliple, ks.rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED)
    for i, t in enumerate(GOLD["kin_t"]):
        table, _, _ = ks.body_schedule(two, F(t), 1e-3, 1)
        np.testing.assert_allclose(table[0, :, 4:6].T, GOLD["kin_fourier_disp_dot"][i], rtol=5e-6, atol=5e-6)
        np.testing.assert_allclose(table[0, :, 6], GOLD["kin_fourier_norm_rot_dot"][i], rtol=5e-6, atol=5e-6)
        th = GOLD["kin_fourier_norm_rot"][i]
        np.testing.assert_allclose(table[0, :, 0], np.cos(th), rtol=0, atol=2e-6)
        np.testing.assert_allclose(table[0, :, 1], np.sin(th), rtol=0, atol=2e-6)


def test_clock_walls_and_body_update_match_the_reference():
    """boundaries.py:519-542 update_BC (float32 clock, wall functions re-evaluated) and particle_motion.py:38-66."""
    g = _grid()
    wall_u = [lambda t: 0.0, lambda t: 0.0, lambda t: 0.1 * np.sin(F(3.0) * t), lambda t: -0.2 * np.cos(F(2.0) * t)]
    bu = boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, -0.2)), F(0.0), wall_u)
    bv = boundaries.Moving_wall_boundary_conditions(2, ((0.0, 0.0), (0.0, 0.0)), F(0.0), ZERO)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    U = grids.GridVariable(grids.GridArray(t(GOLD["u"]), (1.0, 0.5), g), bu)
    V = grids.GridVariable(grids.GridArray(t(GOLD["v"]), (0.5, 1.0), g), bv)
    P = grids.GridVariable(grids.GridArray(t(GOLD["p"]), (0.5, 0.5), g), boundaries.get_pressure_bc_from_velocity((U, V)))
    av = pc.All_Variables(None, (U, V), P, [], 0, None)
    for n in range(len(GOLD["clock_stamps"])):
        av = boundaries.update_BC(av, 5.0e-4)
        assert F(av.velocity[0].bc.time_stamp) == GOLD["clock_stamps"][n]  # bit for bit
        np.testing.assert_allclose(np.asarray(av.velocity[0].bc.bc_values[1], dtype=F), GOLD["clock_wall_u"][n], rtol=1e-6, atol=1e-7)
    x0, y0 = GOLD["x0"], GOLD["y0"]
    bc = boundaries.new_periodic_boundary_conditions(ndim=2, bc_vals=((0.0, 0.0), (0.0, 0.0)), bc_fn=ZERO, time_stamp=F(GOLD["ib_t0"]))
    U = grids.GridVariable(grids.GridArray(t(GOLD["u"]), (1.0, 0.5), g), bc)
    V = grids.GridVariable(grids.GridArray(t(GOLD["v"]), (0.5, 1.0), g), bc)
    one = pc.particle(GOLD["ib_centers"][:1].astype(F), np.array([[0.5, 0.06]], F), GOLD["ib_dparams"][:1].tolist(), GOLD["ib_rparams"][:1].tolist(),
                      pc.Grid1d(len(x0), domain=(0, 2 * np.pi)), lambda gp, th: (x0, y0), ks.displacement, ks.rotation)
    moved = particle_motion.Update_particle_position_Multiple(pc.All_Variables(one, (U, V), P, [], 3, None), float(GOLD["dt"]))
    np.testing.assert_allclose(np.asarray(moved.particles.particle_center), GOLD["ib_new_centers"], rtol=0, atol=2e-7)
    assert moved.Step_count == int(GOLD["ib_step_count"])


def test_domain_integral_matches_the_reference():
    """IBM_Force.py:5-18 Integrate_Field_Fluid_Domain: nested trapezoid rule (last axis with dx, then dy, as written)."""
    from jax_ib_b200 import IBM_Force
    U, _, _ = _variables("periodic")
    got = float(IBM_Force.integrate_trapz(U.data, U.grid.step[0], U.grid.step[1]))
    assert got == pytest.approx(float(GOLD["trapz_u"]), rel=2e-5, abs=1e-6)


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_projection_with_an_opaque_solver_equals_the_reference_bit_for_bit(tag):
    """pressure.py:58-91 in the plug-in fallback (a user `solve` the probes do not recognise): q from the caller, then
    p + q, v - grad q and impose_bc composed from the torch mirror.  Feeding the reference's own q isolates that glue."""
    from jax_ib_b200 import pressure
    U, V, P = _variables(tag)
    q = torch.from_numpy(np.ascontiguousarray(GOLD[f"{tag}_q"]))
    solve = lambda v, q0: grids.GridArray(q, U.grid.cell_center, U.grid)
    new = pressure.projection_and_update_pressure(pc.All_Variables(None, (U, V), P, [], 0, None), solve)
    np.testing.assert_array_equal(new.velocity[0].data.numpy(), GOLD[f"{tag}_proj_u"])
    np.testing.assert_array_equal(new.velocity[1].data.numpy(), GOLD[f"{tag}_proj_v"])
    np.testing.assert_array_equal(new.pressure.data.numpy(), GOLD[f"{tag}_proj_p"])
