"""The oracle against vectors produced by the REFERENCE'S OWN SOURCE FILES (tests/golden/reference_pinned.npz, made by
tests/golden/make_reference_golden.py in the build container: hashimmg/jax_IB's unmodified modules executed on a
NumPy-backed stand-in for jax, oracle/jax_standin).  CPU only; /root/reference is not needed here.

Pins SURVEY 8a rows a1-a19.  Array arithmetic (stencils, ghost cells, differences, the channel and far-field pressure
solves) is compared BIT FOR BIT: both sides are float32 NumPy with the reference's operation order.  Paths through
exp / sin / cos carry a 1e-6 bar (library transcendental functions), the complex-step derivatives standing in for
jax.jacrev 2e-6, the assembled step 2e-6 after one step and 1e-5 after six (measured: 1.5e-7 .. 6e-7).
"""
import os

import numpy as np
import pytest

from oracle import field, ib, kinematics as okin, stencils, stepper

from _adapter import rel_l2

F = np.float32
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pinned.npz"))


def _grid():
    return field.Grid(tuple(int(x) for x in GOLD["shape"]), tuple(tuple(float(y) for y in x) for x in GOLD["domain"]))


def _velocity(tag):
    g = _grid()
    if tag == "periodic":
        bcu = bcv = field.periodic_bc()
    elif tag == "far_field":
        wu, wv = GOLD["far_walls"]
        bcu = field.far_field_bc(tuple(tuple(float(x) for x in r) for r in wu))
        bcv = field.far_field_bc(tuple(tuple(float(x) for x in r) for r in wv))
    else:
        (ulo, uhi), (vlo, vhi) = GOLD["walls"]
        bcu = field.channel_bc(((0.0, 0.0), (float(ulo), float(uhi))))
        bcv = field.channel_bc(((0.0, 0.0), (float(vlo), float(vhi))))
    U = field.Var(GOLD["u"], (1.0, 0.5), g, bcu)
    V = field.Var(GOLD["v"], (0.5, 1.0), g, bcv)
    P = field.Var(GOLD["p"], (0.5, 0.5), g, field.pressure_bc_from_velocity((U, V)))
    return U, V, P


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_stencils_and_ghost_cells_equal_the_reference_bit_for_bit(tag):
    """advection.py:120-333, finite_differences.py:87-146, grids.py shift / boundaries.py:116-273 pad, trim."""
    U, V, P = _velocity(tag)
    vel = (U, V)
    dt = float(GOLD["dt"])
    for k, c in enumerate(vel):
        np.testing.assert_array_equal(stencils.advect_upwind(c, vel, dt), GOLD[f"{tag}_upwind_{k}"])
        np.testing.assert_array_equal(stencils.advect_van_leer(c, vel, dt), GOLD[f"{tag}_van_leer_{k}"])
        np.testing.assert_array_equal(stencils.advect_van_leer_using_limiters(c, vel, dt), GOLD[f"{tag}_van_leer_limiters_{k}"])
        np.testing.assert_array_equal(stencils.laplacian(c), GOLD[f"{tag}_laplacian_{k}"])
        for axis in (0, 1):
            for s in (-2, -1, 1, 2):
                np.testing.assert_array_equal(c.shift(s, axis).data, GOLD[f"{tag}_shift_{k}_{axis}_{s}"])
    np.testing.assert_array_equal(stencils.divergence(vel), GOLD[f"{tag}_divergence"])
    for axis in (0, 1):
        np.testing.assert_array_equal(stencils.forward_difference(P, axis), GOLD[f"{tag}_grad_p_{axis}"])
        np.testing.assert_array_equal(stencils.backward_difference(U, axis), GOLD[f"{tag}_back_u_{axis}"])


def test_delta_and_interpolation_match_the_reference():
    """convolution_functions.py:5-37."""
    got = ib.gaussian_delta(GOLD["delta_x"], GOLD["delta_x0"], float(GOLD["delta_w"]))
    np.testing.assert_allclose(got, GOLD["delta_out"], rtol=1e-6)
    U, V, _ = _velocity("periodic")
    np.testing.assert_allclose(ib.interpolate_dense(U, GOLD["surf_xp"], GOLD["surf_yp"]), GOLD["surf_u"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(ib.interpolate_dense(V, GOLD["surf_xp"], GOLD["surf_yp"]), GOLD["surf_v"], rtol=2e-6, atol=1e-6)


def test_kinematic_laws_and_their_derivatives_match_the_reference():
    """kinematics.py:5-105; the derivatives are what jax.jacrev returns in IBM_Force.py:101-102."""
    disp_param, rot_param = [[2.8, 0.25]], [[np.pi / 2, np.pi / 4, 0.25, 0.3]]
    four = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
    for i, t in enumerate(GOLD["kin_t"]):
        np.testing.assert_allclose(okin.displacement(disp_param, t), GOLD["kin_disp"][i], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(okin.rotation(rot_param, t), GOLD["kin_rot"][i], rtol=1e-6)
        np.testing.assert_allclose(okin.displacement_dot(disp_param, t), GOLD["kin_disp_dot"][i], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(okin.rotation_dot(rot_param, t), GOLD["kin_rot_dot"][i], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(okin.displacement_fourier(four, t), GOLD["kin_fourier_disp"][i], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(okin.rotation_fourier(four, t), GOLD["kin_fourier_rot"][i], rtol=2e-6, atol=2e-6)


def _bodies(n):
    x0, y0 = GOLD["x0"], GOLD["y0"]
    return ib.Bodies(GOLD["ib_centers"][:n].astype(F), [[0.5, 0.06]] * n, GOLD["ib_dparams"][:n].tolist(), GOLD["ib_rparams"][:n].tolist(),
                     lambda g: (x0, y0), okin.DISPLACEMENT["harmonic"], okin.ROTATION["harmonic"])


def test_direct_forcing_and_spreading_of_two_bodies_match_the_reference():
    """IBM_Force.py:20-115 calc_IBM_force_NEW_MULTIPLE: marker kinematics, interpolation, (U_body - U) / dt, arc
    length, dense spreading, sum over bodies -- on both velocity components."""
    g = _grid()
    t0 = F(GOLD["ib_t0"])
    U = field.Var(GOLD["u"], (1.0, 0.5), g, field.periodic_bc(t0))
    V = field.Var(GOLD["v"], (0.5, 1.0), g, field.periodic_bc(t0))
    fx, fy = ib.ibm_force((U, V), _bodies(2), float(GOLD["dt"]), R=None)
    assert rel_l2(fx, GOLD["ib_fx"]) < 1e-6 and rel_l2(fy, GOLD["ib_fy"]) < 1e-6
    # and the windowed sums the CUDA kernels are compared with agree with the reference's dense ones up to the
    # taps the window leaves out (beyond 5 cells: below exp(-12.5) = 3.7e-6 of the peak).  The golden grid starts
    # at (0.5, 0.25), so this also checks that the window follows the domain origin.
    fxw, fyw = ib.ibm_force((U, V), _bodies(2), float(GOLD["dt"]), R=6)
    assert rel_l2(fxw, GOLD["ib_fx"]) < 5e-6 and rel_l2(fyw, GOLD["ib_fy"]) < 5e-6


def test_body_update_matches_the_reference():
    """particle_motion.py:38-66 Update_particle_position_Multiple."""
    g = _grid()
    t0 = F(GOLD["ib_t0"])
    U = field.Var(GOLD["u"], (1.0, 0.5), g, field.periodic_bc(t0))
    V = field.Var(GOLD["v"], (0.5, 1.0), g, field.periodic_bc(t0))
    P = field.Var(GOLD["p"], (0.5, 0.5), g, field.periodic_bc())
    st = stepper.State(_bodies(1), (U, V), P, Step_count=3)
    st = stepper.update_positions(st, float(GOLD["dt"]))
    np.testing.assert_allclose(st.particles.centers, GOLD["ib_new_centers"], rtol=0, atol=2e-7)
    assert st.Step_count == int(GOLD["ib_step_count"])


def test_clock_and_wall_values_match_the_reference():
    """boundaries.py:519-542 update_BC: the float32 clock accumulates dt step by step and the wall functions are
    re-evaluated at the new time stamp."""
    g = _grid()
    wall_u = [lambda t: 0.0, lambda t: 0.0, lambda t: 0.1 * np.sin(F(3.0) * t), lambda t: -0.2 * np.cos(F(2.0) * t)]
    wall_v = [lambda t: 0.0] * 4
    U = field.Var(GOLD["u"], (1.0, 0.5), g, field.channel_bc(((0.0, 0.0), (0.0, -0.2)), F(0.0), wall_u))
    V = field.Var(GOLD["v"], (0.5, 1.0), g, field.channel_bc(((0.0, 0.0), (0.0, 0.0)), F(0.0), wall_v))
    P = field.Var(GOLD["p"], (0.5, 0.5), g, field.periodic_bc())
    st = stepper.State(None, (U, V), P)
    for n in range(len(GOLD["clock_stamps"])):
        st = stepper.update_bc(st, 5.0e-4)
        assert F(st.velocity[0].bc.time_stamp) == GOLD["clock_stamps"][n]          # bit for bit
        np.testing.assert_allclose(np.asarray(st.velocity[0].bc.values[1], dtype=F), GOLD["clock_wall_u"][n], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------- a14-a17, a19 (generator part 2)
def _bcs(tag):
    if tag == "periodic":
        return field.periodic_bc(), field.periodic_bc()
    if tag == "channel":
        (ulo, uhi), (vlo, vhi) = GOLD["walls"]
        return field.channel_bc(((0.0, 0.0), (float(ulo), float(uhi)))), field.channel_bc(((0.0, 0.0), (float(vlo), float(vhi))))
    wu, wv = GOLD["far_walls"]
    return (field.far_field_bc(tuple(tuple(float(x) for x in r) for r in wu)),
            field.far_field_bc(tuple(tuple(float(x) for x in r) for r in wv)))


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_pressure_solve_and_projection_match_the_reference(tag):
    """pressure.py:58-154 + array_utils.py laplacian matrices, run from the reference's files: the right-hand side, the
    1-D operators with their boundary rows, q, p_new = q + p_old, v - grad q, and the wall values re-imposed.  The
    eigen-solve they call (jax_cfd's pseudoinverse) is a restatement on both sides, so what this pins is everything
    AROUND it; a wrong operator row, sign or boundary rule shows up at O(1), far above the 1e-5 bar (float32 round-off
    of two differently ordered transforms of white noise)."""
    from oracle import poisson
    g = _grid()
    bcu, bcv = _bcs(tag)
    U = field.Var(GOLD["u"], (1.0, 0.5), g, bcu)
    V = field.Var(GOLD["v"], (0.5, 1.0), g, bcv)
    P = field.Var(GOLD["p"], (0.5, 0.5), g, field.pressure_bc_from_velocity((U, V)))
    q = poisson.SOLVERS[tag]((U, V))
    assert rel_l2(q, GOLD[f"{tag}_q"]) < 1e-5
    new_v, new_p = poisson.projection_and_update_pressure((U, V), P, tag)
    assert rel_l2(new_v[0].data, GOLD[f"{tag}_proj_u"]) < 2e-6
    assert rel_l2(new_v[1].data, GOLD[f"{tag}_proj_v"]) < 2e-6
    assert rel_l2(new_p.data, GOLD[f"{tag}_proj_p"]) < 2e-6
    if tag != "periodic":  # impose_bc: the wall-aligned rows / columns hold the wall values exactly
        np.testing.assert_array_equal(new_v[1].data[:, -1], GOLD[f"{tag}_proj_v"][:, -1])
    if tag == "far_field":
        np.testing.assert_array_equal(new_v[0].data[-1, :], GOLD[f"{tag}_proj_u"][-1, :])


STEP_CASES = {
    # tag: (convect, wall functions of u, wall values of u at t = 0, wall functions of v, constant force)
    "periodic": ("upwind", None, None, None),
    "channel": ("van_leer_limiters", [lambda t: 0.0, lambda t: 0.0, lambda t: 0.1 * np.sin(F(3.0) * t), lambda t: -0.2 * np.cos(F(2.0) * t)],
                ((0.0, 0.0), (0.0, -0.2)), (0.4, -0.1)),
    "far_field": ("upwind", [lambda t: 0.2] * 4, ((0.2, 0.2), (0.2, 0.2)), None),
}


@pytest.mark.parametrize("tag", ["periodic", "channel", "far_field"])
def test_assembled_step_matches_the_reference(tag):
    """equations.py:194-242 semi_implicit_navier_stokes_timeBC driven by time_stepping.py:109-256 navier_stokes_rk_updated
    (forward Euler), exactly as Flapping_Demo.ipynb cell 9 builds it, with one flapping body: explicit terms, the stored
    pressure's gradient, direct forcing at the predicted velocity, projection, clock and wall update, body update --
    after ONE step (the order of the stages) and after six."""
    convect, fns, vals, force = STEP_CASES[tag]
    g = _grid()
    zero = [lambda t: 0.0] * 4
    make = {"periodic": lambda v, f: field.periodic_bc(F(0.0), f), "channel": lambda v, f: field.channel_bc(v, F(0.0), f),
            "far_field": lambda v, f: field.far_field_bc(v, F(0.0), f)}[tag]
    bcu = make(vals, fns if fns is not None else zero)
    bcv = make(((0.0, 0.0), (0.0, 0.0)), zero)
    U = field.Var(GOLD["step_u0"], (1.0, 0.5), g, bcu)
    V = field.Var(GOLD["step_v0"], (0.5, 1.0), g, bcv)
    P = field.Var(np.zeros(g.shape, F), (0.5, 0.5), g, field.pressure_bc_from_velocity((U, V)))
    forcing = None if force is None else (lambda vv: tuple(F(f) * np.ones_like(x.data) for f, x in zip(force, vv)))
    step = stepper.make_step(float(GOLD["dt"]), float(GOLD["step_viscosity"]), float(GOLD["step_density"]), convect, tag, forcing)
    st = stepper.State(_bodies(1), (U, V), P)
    for n in range(int(GOLD["step_n"])):
        st = step(st)
        if n == 0:
            assert rel_l2(st.velocity[0].data, GOLD[f"step1_{tag}_u"]) < 2e-6
            assert rel_l2(st.velocity[1].data, GOLD[f"step1_{tag}_v"]) < 2e-6
            assert rel_l2(st.pressure.data, GOLD[f"step1_{tag}_p"]) < 2e-5
    assert rel_l2(st.velocity[0].data, GOLD[f"step_{tag}_u"]) < 1e-5
    assert rel_l2(st.velocity[1].data, GOLD[f"step_{tag}_v"]) < 1e-5
    assert rel_l2(st.pressure.data, GOLD[f"step_{tag}_p"]) < 1e-4
    np.testing.assert_allclose(st.particles.centers, GOLD[f"step_{tag}_centers"], rtol=0, atol=3e-7)
    assert F(st.velocity[0].bc.time_stamp) == GOLD[f"step_{tag}_time"]
    assert st.Step_count == int(GOLD[f"step_{tag}_count"])


# ------------------------------------------------------------------------------- a21, a22 (generator part 3)
def test_tracer_coupling_matches_the_reference():
    """MD/CFD_MD_coupling.py:7-27 (fractional index x/dx - offset, bilinear, taps outside the array are zero) and
    MD/simulate.py:81-97 at kT = 0 with the notebooks' masked shift: three Brownian steps in a frozen velocity field."""
    from oracle import tracers
    U, V, _ = _velocity("periodic")
    pts = GOLD["md_pts"]
    np.testing.assert_array_equal(tracers.cfd_force((U, V), pts), GOLD["md_force"])  # bit for bit, edge and outside points too
    st = tracers.BrownianState(pts.copy(), 1.0)
    for _ in range(3):
        st = tracers.brownian_step(st, (U, V), float(GOLD["dt"]), gamma=0.8, kT=0.0, is_mobile=GOLD["md_mobile"])
    np.testing.assert_allclose(st.position, GOLD["md_pos3"], rtol=0, atol=3e-7)
    np.testing.assert_array_equal(st.position[~GOLD["md_mobile"]], pts[~GOLD["md_mobile"]])
    got = tracers.harmonic_morse(GOLD["morse_dr"], F(5.0), F(0.7), F(10.0), F(0.9), F(1.0))
    np.testing.assert_allclose(got, GOLD["morse_u"], rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["ellipse", "circle", "rose"])
def test_penalty_mask_matches_the_reference(name):
    """penalty/util_funs.py:26-93 project_particle / calc_perm with the shapes of penalty/parameteric_fns.py: polar
    angle (three-branch arctan), segment lookup, linear R(theta), level set, smoothing."""
    from oracle import penalty
    g = _grid()
    theta = np.linspace(0, 2 * np.pi, 61).astype(F)
    shape_fn = {"ellipse": lambda: penalty.param_ellipse([0.45, 0.2], theta), "circle": lambda: penalty.param_circle([0.3], theta),
                "rose": lambda: penalty.param_rose([0.4, 2.0], theta)}[name]
    Rth = shape_fn()
    np.testing.assert_allclose(Rth, GOLD[f"pen_R_{name}"], rtol=2e-6, atol=1e-7)
    Rth = GOLD[f"pen_R_{name}"]  # same table on both sides from here on
    with np.errstate(all="ignore"):
        Rfinal, _ = penalty.polar_radius(g, (1.6, 1.0), Rth)
        perm = penalty.calc_perm(g, (1.6, 1.0), Rth, penalty.tanh_smoothing, tuple(float(x) for x in GOLD["pen_know"]))
    ok = np.isfinite(GOLD[f"pen_Rfinal_{name}"])
    assert np.array_equal(ok, np.isfinite(Rfinal))
    np.testing.assert_array_equal(Rfinal[ok], GOLD[f"pen_Rfinal_{name}"][ok])  # same NumPy operations in the same order
    np.testing.assert_array_equal(perm, GOLD[f"pen_perm_{name}"])


def test_penalty_forcing_and_stepper_match_the_reference():
    """penalty/util_funs.py:6-16 arbitrary_obstacle and equations.py:245-280 semi_implicit_navier_stokes_penalty with
    time_stepping.py:259-376 forward_euler_penalty: three steps between channel walls."""
    from oracle import penalty
    U, V, _ = _velocity("periodic")
    perm = GOLD["pen_perm_ellipse"]
    fu, fv = penalty.obstacle_forcing((0.3, -0.1), perm)((U, V))
    np.testing.assert_allclose(fu, GOLD["pen_force_u"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(fv, GOLD["pen_force_v"], rtol=1e-6, atol=1e-6)
    g = _grid()
    zero = [lambda t: 0.0] * 4
    bc = field.channel_bc(((0.0, 0.0), (0.0, 0.0)), F(0.0), zero)
    U = field.Var(GOLD["step_u0"], (1.0, 0.5), g, bc)
    V = field.Var(GOLD["step_v0"], (0.5, 1.0), g, bc)
    P = field.Var(np.zeros(g.shape, F), (0.5, 0.5), g, field.pressure_bc_from_velocity((U, V)))
    step = stepper.make_step(float(GOLD["dt"]), float(GOLD["step_viscosity"]), float(GOLD["step_density"]), "upwind", "channel",
                             penalty.obstacle_forcing((0.3, -0.1), perm), penalty=True)
    st = stepper.State(None, (U, V), P)
    for _ in range(3):
        st = step(st)
    assert rel_l2(st.velocity[0].data, GOLD["pen_step_u"]) < 5e-6
    assert rel_l2(st.velocity[1].data, GOLD["pen_step_v"]) < 5e-6
    assert rel_l2(st.pressure.data, GOLD["pen_step_p"]) < 5e-5
    assert F(st.velocity[0].bc.time_stamp) == GOLD["pen_step_time"]


# ------------------------------------------------------------------------------- multi-body Fourier laws (generator part 4)
FOUR_D = [[0.2, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.0], [-0.1, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], 0.5]]
FOUR_R = [[0.8, 0.3, 0.1, [0.5, 0.25, 0.1], [0.05, 0.02, 0.3], 1.3, 1.0], [0.5, 0.15, 0.7, [0.3, 0.0, 0.2], [0.1, 0.2, 0.0], -0.7, 0.5]]


def test_two_bodies_on_the_journal_bearing_laws_match_the_reference():
    """kinematics.py:14-105 (Fourier displacement, NORMALISED Fourier rotation) with their jacrev derivatives, the
    per-body loop IBM_Force.py:89-106, and three assembled steps in which particle_motion.py:38-66 moves BOTH centres
    with the multi-body law evaluated on all parameter sets at once."""
    for i, t in enumerate(GOLD["kin_t"]):
        np.testing.assert_allclose(okin.rotation_fourier_normalized(FOUR_R, t), GOLD["kin_fourier_norm_rot"][i], rtol=3e-6, atol=1e-6)
        np.testing.assert_allclose(okin.displacement_fourier_dot(FOUR_D, t), GOLD["kin_fourier_disp_dot"][i], rtol=5e-6, atol=5e-6)
        np.testing.assert_allclose(okin.rotation_fourier_normalized_dot(FOUR_R, t), GOLD["kin_fourier_norm_rot_dot"][i], rtol=5e-6, atol=5e-6)
    g = _grid()
    x0, y0 = GOLD["x0"], GOLD["y0"]
    bodies = ib.Bodies(GOLD["ib_centers"].astype(F), [[0.5, 0.06]] * 2, FOUR_D, FOUR_R, lambda gp: (x0, y0),
                       okin.DISPLACEMENT["fourier"], okin.ROTATION["fourier_normalized"])
    bc = field.periodic_bc(F(0.37), [lambda t: 0.0] * 4)
    U = field.Var(GOLD["step_u0"], (1.0, 0.5), g, bc)
    V = field.Var(GOLD["step_v0"], (0.5, 1.0), g, bc)
    P = field.Var(np.zeros(g.shape, F), (0.5, 0.5), g, field.periodic_bc())
    fx, fy = ib.ibm_force((U, V), bodies, float(GOLD["dt"]), R=None)
    assert rel_l2(fx, GOLD["two_fx"]) < 2e-6 and rel_l2(fy, GOLD["two_fy"]) < 2e-6
    step = stepper.make_step(float(GOLD["dt"]), float(GOLD["step_viscosity"]), float(GOLD["step_density"]), "upwind", "periodic")
    st = stepper.State(bodies, (U, V), P)
    for _ in range(3):
        st = step(st)
    assert rel_l2(st.velocity[0].data, GOLD["two_step_u"]) < 5e-6
    assert rel_l2(st.velocity[1].data, GOLD["two_step_v"]) < 5e-6
    assert rel_l2(st.pressure.data, GOLD["two_step_p"]) < 5e-5
    np.testing.assert_allclose(st.particles.centers, GOLD["two_step_centers"], rtol=0, atol=5e-7)


@pytest.mark.skipif(not os.path.isdir("/root/reference/jax_ib"), reason="the reference tree only exists in the build container")
def test_committed_fixture_is_what_the_reference_files_produce_today():
    """Regenerates every golden array from /root/reference in a child process (the stand-in shadows `jax` on sys.path,
    which must not leak into this process) and compares with the committed .npz, bit for bit."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        code = ("import sys, numpy as np; sys.path.insert(0, %r); import make_reference_golden as m; "
                "np.savez(%r, **m.generate())" % (os.path.join(root, "tests", "golden"), os.path.join(tmp, "fresh.npz")))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        fresh = np.load(os.path.join(tmp, "fresh.npz"))
        assert sorted(fresh.files) == sorted(GOLD.files)
        for k in GOLD.files:
            np.testing.assert_array_equal(fresh[k], GOLD[k], err_msg=k)


@pytest.mark.parametrize("name,oracle_fixture", [("flap600", "flap600_20.npz"), ("bearing300", "bearing300_20.npz")])
def test_oracle_trajectory_fixtures_equal_the_reference_notebook_runs(name, oracle_fixture):
    """tests/golden/reference_notebook_<name>.npz are the reference's NOTEBOOKS themselves -- Flapping_Demo.ipynb (600 x 600,
    one 298-marker ellipse) and journal_bearing_demo.ipynb (300 x 300, two concentric circles on the normalised Fourier
    laws, passive tracers) with their code cells executed verbatim on the stand-in for 20 steps, dense marker sums
    (make_reference_notebook_golden.py).  tests/golden/<oracle_fixture> is the same run from the oracle, the fixture the
    GPU tests have replayed since round 1.  Comparing the two ties every CUDA-vs-golden assertion to the notebooks' own
    output.  Bars: 2e-6 for the velocities; for the pressure twice the float32 round-off floor the oracle fixture
    records for its trajectory (|p| << |u| in the bearing run: floor 7e-5)."""
    here = os.path.dirname(os.path.abspath(__file__))
    nb = np.load(os.path.join(here, "golden", f"reference_notebook_{name}.npz"))
    gold = np.load(os.path.join(here, "golden", oracle_fixture))
    assert int(nb["steps"]) == int(gold["step_count"]) == 20 and list(nb["snapshot_steps"]) == [0, 10]
    assert np.array_equal(nb["crop"], gold["crop"])
    for key, bar in (("u", 2e-6), ("v", 2e-6), ("p", 2.0 * float(gold["p_noise"]))):
        assert rel_l2(gold[key + "_crop"], nb[key + "_crop"]) < bar, key
        assert float(gold[key + "_l2"]) == pytest.approx(float(nb[key + "_l2"]), rel=1e-6 if key != "p" else 1e-4)
    np.testing.assert_array_equal(gold["centers"], nb["centers"])
    assert F(gold["time"]) == F(nb["time"])


def test_tracers_in_the_journal_bearing_notebook_run():
    """particle_motion.py:6-35 Update_particle_position_Multiple_and_MD_Step inside the assembled step: WHICH velocity
    field and WHICH body state the Brownian sub-step sees (time_stepping.py:253 runs it last, on the projected field).
    The notebook's 1401 tracers were advanced for 20 steps by the reference's own files (passive: the jax_md pair
    potential is not restated, so there is no wall repulsion on either side); the oracle repeats the run from the same
    start positions."""
    from jax_ib_b200 import configs
    from oracle import tracers
    from _adapter import oracle_state, oracle_step
    here = os.path.dirname(os.path.abspath(__file__))
    nb = np.load(os.path.join(here, "golden", "reference_notebook_bearing300.npz"))
    cfg = configs.bearing300()
    st = oracle_state(cfg)
    st.MD_var = tracers.BrownianState(nb["md_pos0"].copy())
    md = lambda s: tracers.brownian_step(s.MD_var, s.velocity, cfg["dt"], float(nb["md_gamma"]), 0.0, None, nb["md_mobile"])
    step = oracle_step(cfg, ib_window=6, md_step=md)
    for _ in range(int(nb["steps"])):
        st = step(st)
    moved = np.abs(nb["md_pos"] - nb["md_pos0"]).max()
    assert moved > 5e-3  # the tracers really travelled (|u| dt steps ~ 4e-4 each)
    np.testing.assert_allclose(st.MD_var.position, nb["md_pos"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(st.MD_var.position[~nb["md_mobile"]], nb["md_pos0"][~nb["md_mobile"]])
    a, b, c, d = (int(x) for x in nb["crop"])
    assert rel_l2(st.velocity[0].data[a:b, c:d], nb["u_crop"]) < 2e-6


def test_bench_workload_three_steps_by_the_reference_files():
    """ib2048 -- the configuration bench.py times -- advanced three steps by the reference's own files with dense marker
    sums (tests/golden/make_reference_workload_golden.py, ~10 min on NumPy) against the windowed oracle the GPU test of the
    same workload uses (tests/test_gpu_round2.py::test_bench_workloads_match_windowed_oracle): crops around the four
    bodies, 64 x 64 block means of the whole field, norms, centres, clock."""
    from jax_ib_b200 import configs
    from _adapter import oracle_state, oracle_step
    here = os.path.dirname(os.path.abspath(__file__))
    ref = np.load(os.path.join(here, "golden", "reference_ib2048_3.npz"))
    cfg = configs.ib_periodic(2048, 2)
    st = oracle_state(cfg)
    step = oracle_step(cfg, ib_window=6)
    for _ in range(int(ref["steps"])):
        st = step(st)
    for key, var in (("u", st.velocity[0]), ("v", st.velocity[1]), ("p", st.pressure)):
        full = var.data
        crops = np.stack([full[a:b, c:d] for a, b, c, d in ref["boxes"]])
        assert rel_l2(crops, ref[key + "_crops"]) < 2e-6, key
        blocks = full.astype(np.float64).reshape(64, 32, 64, 32).mean(axis=(1, 3))
        assert rel_l2(blocks, ref[key + "_blocks"]) < 2e-6, key
        assert float(np.sqrt((full.astype(np.float64) ** 2).sum())) == pytest.approx(float(ref[key + "_l2"]), rel=1e-6)
    np.testing.assert_allclose(st.particles.centers, ref["centers"], rtol=0, atol=5e-7)
    assert F(st.velocity[0].bc.time_stamp) == F(ref["time"])
    assert st.Step_count == int(ref["steps"])


def test_channel_trajectory_fixture_is_bit_identical_to_the_reference_run():
    """G3's parity case (1024 x 256 channel, van Leer, body force, Poiseuille + noise, five steps): the oracle's fixture
    the GPU test replays (tests/golden/channel1024x256_5.npz) against the same run by the reference's own files with its
    eigh + matmul pressure solve (tests/golden/make_reference_workload_golden.py channel)."""
    here = os.path.dirname(os.path.abspath(__file__))
    ref = np.load(os.path.join(here, "golden", "reference_channel1024x256_5.npz"))
    gold = np.load(os.path.join(here, "golden", "channel1024x256_5.npz"))
    assert int(ref["steps"]) == int(gold["steps"]) == 5 and np.array_equal(ref["crop"], gold["crop"])
    for key in "uvp":
        np.testing.assert_array_equal(gold[key + "_crop"], ref[key + "_crop"])
        assert float(gold[key + "_l2"]) == float(ref[key + "_l2"])
    assert F(gold["time"]) == F(ref["time"])


def test_small_periodic_trajectory_fixture_equals_the_reference_run():
    """small64_50.npz (64 x 64, one flapping ellipse, FIFTY steps -- the longest fixture) against the same run by the
    reference's files (make_reference_workload_golden.py small64): 1e-7 on the velocities after 50 steps, centre and
    float32 clock (0.09999994: fifty float32 additions of 2e-3) identical."""
    here = os.path.dirname(os.path.abspath(__file__))
    ref = np.load(os.path.join(here, "golden", "reference_small64_50.npz"))
    gold = np.load(os.path.join(here, "golden", "small64_50.npz"))
    assert int(ref["steps"]) == int(gold["steps"]) == 50 and np.array_equal(ref["crop"], gold["crop"])
    for key, bar in (("u", 5e-7), ("v", 5e-7), ("p", 2.0 * float(gold["p_noise"]))):
        assert rel_l2(gold[key + "_crop"], ref[key + "_crop"]) < bar, key
    np.testing.assert_array_equal(gold["centers"], ref["centers"])
    assert F(gold["time"]) == F(ref["time"])
