"""CPU-side checks (no GPU needed): the C-ABI library builds, loads and exports every symbol the
header declares; the host-side mirror (kinematics table, ghost-cell shift, clock) agrees with the
oracle; the golden fixtures are what the oracle produces."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from jax_ib_b200 import _lib, boundaries, build, configs, demo, grids, kinematics
from oracle import field, ib, kinematics as okin

from _adapter import oracle_state, oracle_step, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def _header_functions():
    text = open(os.path.join(ROOT, "include", "jaxib_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jaxib_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jaxib_b200.h but not exported"
    # and the Python binding covers the same set
    assert set(names) == set(_lib.exported_symbols())
    l = _lib.load()
    assert l.jaxib_compiled_arch() == 100
    assert b"sm_100a" in l.jaxib_version()


def test_argument_errors_without_a_gpu():
    l = _lib.load()
    assert l.jaxib_plan_create(None, None) == -1
    g = _lib.Grid()
    g.nx, g.ny, g.batch, g.dx, g.dy = 0, 8, 1, 1.0, 1.0
    h = ctypes.c_void_p()
    assert l.jaxib_plan_create(ctypes.byref(g), ctypes.byref(h)) == -1
    assert b"nx" in l.jaxib_last_error()
    assert l.jaxib_step(None, None, None, None, None, None, None, 1, 1, None, None, None, None, 0) == -1


def test_plain_c_consumer_of_the_header(tmp_path):
    """tests/c_abi_smoke.c: gcc (no CUDA, no C++) against include/jaxib_b200.h -- every entry point is resolved
    through the header's own prototypes and the argument-error paths run without a GPU.  Header drift fails here."""
    import subprocess
    exe = str(tmp_path / "c_abi_smoke")
    src = os.path.join(ROOT, "tests", "c_abi_smoke.c")
    cc = subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-ldl", "-o", exe],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe, build.build()], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "26 entry points" in run.stdout
    # every function the header declares is exercised by the C consumer
    text = open(src).read()
    for name in _header_functions():
        assert f"LOAD({name})" in text, f"{name} is declared in the header but not loaded by c_abi_smoke.c"


def test_xla_ffi_shim_type_checks_against_the_api_shaped_header():
    """jaxlib is not installable here, so the XLA-FFI shim cannot be built for real; it is type-checked (g++
    -fsyntax-only) against tests/xla_ffi_stub/xla/ffi/api/ffi.h, an API-shaped stand-in whose handler macro asserts
    that every implementation is invocable with exactly the arguments its binding declares -- and the shim must
    wrap every stage entry point of the C ABI."""
    import subprocess
    shim = os.path.join(ROOT, "jax_ib_b200", "csrc", "xla_ffi_shim.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "xla_ffi_stub"),
                        "-I/usr/local/cuda/include", shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    text = open(shim).read()
    not_wrapped = {"jaxib_version", "jaxib_last_error", "jaxib_compiled_arch", "jaxib_plan_create", "jaxib_plan_destroy",
                   "jaxib_scratch_bytes", "jaxib_ib_marker_force", "jaxib_ib_spread_bodies", "jaxib_poisson_rows_fwd",
                   "jaxib_poisson_cols", "jaxib_poisson_rows_inv", "jaxib_slab_step", "jaxib_step_profile"}
    for name in _header_functions():
        if name not in not_wrapped:
            assert name + "(" in text, f"{name} has no XLA-FFI handler"
    # a mismatched handler is rejected by the stand-in (the check is not vacuous)
    bad = os.path.join(os.path.dirname(shim), "..", "..", "tests", "xla_ffi_stub", "negative_example.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "xla_ffi_stub"),
                        "-I/usr/local/cuda/include", bad], capture_output=True, text=True)
    assert r.returncode != 0 and "not invocable" in r.stderr


def test_sass_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {build.build()} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.parametrize("cfg_fn", [configs.flap600, configs.bearing300])
def test_body_schedule_matches_oracle_kinematics(cfg_fn):
    cfg = cfg_fn()
    particles = demo.make_particles(cfg)
    n = 40
    table, times, centers = kinematics.body_schedule(particles, 0.0, cfg["dt"], n)
    st = oracle_state(cfg)
    step_t = F(0.0)
    ob = st.particles
    for k in range(n):
        for b in range(len(ob.centers)):
            th = np.asarray(ob.rotation[0]([ob.rotation_param[b]], step_t), dtype=F).reshape(-1)[0]
            om = np.asarray(ob.rotation[1]([ob.rotation_param[b]], step_t), dtype=F).reshape(-1)[0]
            U0 = np.asarray(ob.displacement[1]([ob.displacement_param[b]], step_t), dtype=F).reshape(2, -1)[:, 0]
            np.testing.assert_allclose(table[k, b, 0], np.cos(th), atol=2e-7)
            np.testing.assert_allclose(table[k, b, 1], np.sin(th), atol=2e-7)
            np.testing.assert_allclose(table[k, b, 4:6], U0, rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(table[k, b, 6], om, rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(table[k, b, 2:4], ob.centers[b], rtol=0, atol=2e-6)
        assert times[k] == step_t
        step_t = F(step_t + F(cfg["dt"]))  # boundaries.py:530
        U = np.asarray(ob.displacement[1](list(ob.displacement_param), step_t), dtype=F).reshape(2, -1)
        c = np.asarray(ob.centers, dtype=F)
        ob = ob.replace_centers(np.stack([c[:, 0] + F(cfg["dt"]) * U[0], c[:, 1] + F(cfg["dt"]) * U[1]], axis=1))
    np.testing.assert_allclose(centers[-1], ob.centers, rtol=0, atol=3e-6)


def test_grid_and_shift_mirror_the_oracle():
    g = grids.Grid((6, 5), domain=((0.0, 3.0), (1.0, 2.0)))
    og = field.Grid((6, 5), ((0.0, 3.0), (1.0, 2.0)))
    assert g.step == og.step and g.cell_faces == og.cell_faces
    for off in (g.cell_center,) + g.cell_faces:
        for a, b in zip(g.axes(off), og.axes(off)):
            np.testing.assert_array_equal(a.numpy(), b)
    data = np.random.default_rng(0).standard_normal((6, 5)).astype(F)
    vals = ((0.0, 0.0), (0.7, -0.4))
    bc = boundaries.Moving_wall_boundary_conditions(2, vals, 0.0, [lambda t: 0.0] * 4)
    obc = field.channel_bc(vals)
    for off in ((1.0, 0.5), (0.5, 1.0)):
        gv = grids.GridVariable(grids.GridArray(torch.from_numpy(data), off, g), bc)
        ov = field.Var(data, off, og, obc)
        for axis in (0, 1):
            for k in (-2, -1, 1, 2):
                got = gv.shift(k, axis)
                want = ov.shift(k, axis)
                np.testing.assert_allclose(got.data.numpy(), want.data, rtol=0, atol=1e-7)
                assert got.offset == tuple(want.offset)
        np.testing.assert_array_equal(gv.impose_bc().data.numpy(), ov.impose_bc().data)
    pbc = boundaries.get_pressure_bc_from_velocity((gv, gv))
    assert pbc.types[0][0] == "periodic" and pbc.types[1] == ("neumann", "neumann")
    with pytest.raises(grids.InconsistentOffsetError):
        grids.GridArray(data, (1.0, 0.5), g) + grids.GridArray(data, (0.5, 1.0), g)


def test_update_bc_advances_the_float32_clock():
    cfg = configs.flap600()
    av = demo.make_all_variables(cfg)
    for _ in range(3):
        av = boundaries.update_BC(av, cfg["dt"])
    t = F(0.0)
    for _ in range(3):
        t = F(t + F(cfg["dt"]))
    assert av.velocity[0].bc.time_stamp == t


def test_golden_small_fixture_is_reproduced_by_the_oracle():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "small64_50.npz"))
    cfg = configs.small_periodic(64)
    st = oracle_state(cfg)
    step = oracle_step(cfg, ib_window=None)
    for _ in range(int(gold["steps"])):
        st = step(st)
    np.testing.assert_array_equal(st.velocity[0].data, gold["u_crop"])
    np.testing.assert_array_equal(st.pressure.data, gold["p_crop"])
    np.testing.assert_array_equal(st.particles.centers, gold["centers"])


def test_product_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        demo.make_stepper(configs.small_periodic(32))


def test_ctypes_structs_match_the_c_header(tmp_path):
    """ABI drift guard: sizes and field offsets of jaxib_grid / jaxib_bodies / jaxib_slab as gcc lays them out
    from include/jaxib_b200.h must equal the ctypes mirrors in jax_ib_b200/_lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probes = {
        "jaxib_grid": ["nx", "dt", "bc", "force", "ib_window", "row0", "own_hi", "nx_global"],
        "jaxib_bodies": ["n_markers", "body_offset", "y0", "max_radius"],
        "jaxib_slab": ["rank", "cpr", "seam_idle", "body_first", "marker_count", "u", "markers", "colbuf", "specbuf", "flags"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "jaxib_b200.h"', "int main(void) {"]
    for st, fields in probes.items():
        lines.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for f in fields:
            lines.append(f'  printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirrors = {"jaxib_grid": _lib.Grid, "jaxib_bodies": _lib.Bodies, "jaxib_slab": _lib.Slab}
    for st, fields in probes.items():
        assert int(out[st]) == ctypes.sizeof(mirrors[st]), st
        for f in fields:
            assert int(out[f"{st}.{f}"]) == getattr(mirrors[st], f).offset, (st, f)
    assert _lib.MAX_PEERS == 16


def test_peer_block_layout_is_aligned_and_disjoint():
    from jax_ib_b200 import slab
    L = slab.SlabLayout(8192, 8192, 8, 3, 16)
    blk = slab.PeerBlock(L, 4768)
    spans = sorted((off, off + blk.size[name], name) for name, off in blk.offset.items())
    for (a0, a1, _), (b0, _, _) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert all(off % 256 == 0 for off in blk.offset.values())
    assert blk.nbytes >= spans[-1][1] and blk.size["flags"] == 4 * slab.PeerBlock.FLAG_WORDS
    assert blk.size["colbuf"] == 8192 * L.cpr * 8 and blk.size["specbuf"] == (L.nloc + 1) * L.sp * 8


# ---- column sweep (csrc/poisson_sweep.cu) restated in NumPy: same operator as the oracle's pseudo-inverse -------------
@pytest.mark.parametrize("shape,slabs", [((96, 64), 1), ((100, 40), 1), ((600, 24), 1), ((96, 64), 3), ((128, 32), 8)])
def test_column_sweep_model_matches_the_oracle_pseudoinverse(shape, slabs):
    """tools/sweep_model.py restates the kernel's arithmetic (16-row segments, 32-lane scan of the carried values,
    cyclic closure, two-prefix-sum mean mode, slab passes A / B).  Its solution must be the oracle's
    `pseudoinverse(..., 'rfft')` (pressure.py:94-108 through jax_cfd fast_diagonalization) to float32 round-off:
    partial last segments (100, 600 rows), several segments per lane (600 rows), 3 and 8 slabs."""
    import importlib.util
    import os
    from oracle import poisson
    spec = importlib.util.spec_from_file_location(
        "sweep_model", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sweep_model.py"))
    sm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sm)
    nx, ny = shape
    dx, dy = 0.05, 0.03
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal(shape)
    rhs -= rhs.mean()
    ops_p = [poisson.laplacian_matrix(nx, dx), poisson.laplacian_matrix(ny, dy)]
    want = poisson.pseudoinverse(ops_p, np.float64, True, "rfft")(rhs)
    got = sm.solve(rhs.astype(np.float32), dx, dy, slabs)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-6
    # and it does solve the discrete Poisson equation: the 5-point Laplacian of the solution is the right-hand side
    lap = ((np.roll(got, 1, 0) - 2 * got + np.roll(got, -1, 0)) / dx ** 2 + (np.roll(got, 1, 1) - 2 * got + np.roll(got, -1, 1)) / dy ** 2)
    assert np.linalg.norm(lap - rhs) / np.linalg.norm(rhs) < 2e-4


def test_ib_window_grows_with_the_cell_aspect_ratio():
    """IBM_Force.py:67 spreads with a radial Gaussian of width dx: the window (in cells) has to cover the same number
    of standard deviations along y when dy < dx; isotropic grids keep the caller's window."""
    from jax_ib_b200 import engine
    assert engine.effective_ib_window(6, (0.025, 0.025)) == 6
    assert engine.effective_ib_window(6, (0.03, 0.05)) == 6
    assert engine.effective_ib_window(6, (0.05, 0.03)) == 10
    assert engine.effective_ib_window(8, (0.1, 0.01)) == 16  # clamped to what the kernels take
