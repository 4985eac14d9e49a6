"""Slab decomposition: index arithmetic and exchanges on CPU (incl. world_size-2 gloo),
and on the GPU the decomposed step against the single-GPU step -- bit for bit."""
import contextlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from jax_ib_b200 import slab  # noqa: E402


# ---- CPU: layout --------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,world,halo", [(64, 64, 1, 16), (64, 32, 2, 16), (256, 128, 4, 16), (8192, 8192, 8, 16)])
def test_layout_covers_every_row_once(nx, ny, world, halo):
    owned = []
    for r in range(world):
        L = slab.SlabLayout(nx, ny, world, r, halo)
        assert L.nxl == L.nloc + 2 * halo
        g = L.global_rows()
        assert list(g[halo:halo + L.nloc]) == list(range(L.own_lo, L.own_hi))
        assert g[0] == (L.own_lo - halo) % nx and g[-1] == (L.own_hi + halo - 1) % nx
        owned += list(range(L.own_lo, L.own_hi))
        seg = L.main_segment()
        # the main segment is exactly the un-wrapped part of the local array
        un = [(L.row0 + i) for i in range(L.nxl) if 0 <= L.row0 + i < nx]
        assert seg.rows == len(un) and seg.global0 == un[0] and L.row0 + seg.start == un[0]
        seam = L.seam_row()
        assert (seam is not None) == (r == 0)
        if seam is not None:
            assert g[seam.start] == nx - 1 == seam.global0
        # column ranges tile the padded spectrum pitch and cover the ny/2 + 4 columns the row kernels write
        assert L.cpr % 8 == 0 and L.sp == world * L.cpr >= ny // 2 + 4
        idx = L.inverse_row_index().reshape(world, L.nloc + 1)
        for d in range(world):
            assert list(idx[d]) == [(d * L.nloc + i) % nx for i in range(L.nloc + 1)]
    assert owned == list(range(nx))
    # markers a slab needs: everything within R + 2 rows of its block; the outer slabs take the off-grid ones
    R = 6
    first, last = slab.SlabLayout(nx, ny, world, 0, halo), slab.SlabLayout(nx, ny, world, world - 1, halo)
    assert first.marker_range(R)[0] < -10 ** 9 and last.marker_range(R)[1] > 10 ** 9
    if world > 1:
        assert first.marker_range(R)[1] == first.own_hi + R and last.marker_range(R)[0] == last.own_lo - R - 2
        seg, lo, hi = first.top_segment(R)
        assert (seg.start, seg.rows, seg.global0) == (0, halo, nx - halo) and lo == nx - R - 3 and hi > 10 ** 9
    else:
        assert first.top_segment(R) is None


def test_layout_rejects_bad_splits():
    with pytest.raises(ValueError):
        slab.SlabLayout(100, 64, 3, 0, 10)
    with pytest.raises(ValueError):
        slab.SlabLayout(64, 64, 8, 0, 16)  # halo deeper than the slab


def _fake_exchange_buffers(world, nloc, cpr, halo, ny, seed=0):
    """Per rank: send [world, nloc, cpr, 2], force [2, M], fields [nloc + 2 halo, ny] x 3."""
    rng = np.random.default_rng(seed)
    out = []
    for r in range(world):
        out.append(dict(
            send=torch.from_numpy(rng.standard_normal((world, nloc, cpr, 2)).astype(np.float32)),
            force=torch.from_numpy(rng.standard_normal((2, 7)).astype(np.float32)),
            fields=[torch.from_numpy(rng.standard_normal((nloc + 2 * halo, ny)).astype(np.float32)) for _ in range(3)],
        ))
    return out


def _halo_views(fields, halo, nloc):
    return ([t[halo:2 * halo] for t in fields], [t[nloc:nloc + halo] for t in fields],
            [t[0:halo] for t in fields], [t[halo + nloc:] for t in fields])


def _local_reference(world, nloc, cpr, halo, ny):
    bufs = _fake_exchange_buffers(world, nloc, cpr, halo, ny)
    comm = slab.LocalComm(world)
    recv = [torch.zeros_like(b["send"]) for b in bufs]
    comm.transpose([b["send"] for b in bufs], recv)
    views = [_halo_views(b["fields"], halo, nloc) for b in bufs]
    comm.halos(*[[v[k] for v in views] for k in range(4)])
    return bufs, recv


def test_local_comm_semantics():
    world, nloc, cpr, halo, ny = 3, 4, 8, 2, 6
    orig = _fake_exchange_buffers(world, nloc, cpr, halo, ny)
    bufs, recv = _local_reference(world, nloc, cpr, halo, ny)
    for d in range(world):
        for r in range(world):
            assert torch.equal(recv[d][r], orig[r]["send"][d])
    for r in range(world):
        nxt, prev = (r + 1) % world, (r - 1) % world
        for k in range(3):
            # upper halo = first owned rows of the next slab, lower halo = last owned rows of the previous one
            assert torch.equal(bufs[r]["fields"][k][halo + nloc:], orig[nxt]["fields"][k][halo:2 * halo])
            assert torch.equal(bufs[r]["fields"][k][:halo], orig[prev]["fields"][k][nloc:nloc + halo])


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nloc, cpr, halo, ny = 4, 8, 2, 6
        mine = _fake_exchange_buffers(world, nloc, cpr, halo, ny)[rank]
        comm = slab.DistComm()
        recv = torch.zeros_like(mine["send"])
        comm.transpose([mine["send"]], [recv])
        views = _halo_views(mine["fields"], halo, nloc)
        comm.halos(*[[v] for v in views])
        q.put((rank, recv.numpy(), mine["force"].numpy(), [f.numpy() for f in mine["fields"]]))
    finally:
        dist.destroy_process_group()


def test_dist_comm_matches_local_comm_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, recv, force, fields = q.get(timeout=120)
        got[r] = (recv, force, fields)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bufs, recv_ref = _local_reference(world, 4, 8, 2, 6)
    for r in range(world):
        np.testing.assert_array_equal(got[r][0], recv_ref[r].numpy())
        for k in range(3):
            np.testing.assert_array_equal(got[r][2][k], bufs[r]["fields"][k].numpy())


# ---- GPU: decomposed step == single-GPU step ------------------------------------------------------
def _seam_cfg(n):
    """Bodies on the periodic seam (markers off the grid), on a slab boundary and in a slab's interior."""
    from jax_ib_b200 import configs
    cfg = configs.ib_periodic(n, 2, dx=0.025)
    L = n * 0.025
    cfg["bodies"]["centers"] = [[0.004 * L, 0.3 * L], [0.5 * L, 0.55 * L], [0.985 * L, 0.7 * L], [0.27 * L, 0.8 * L]]
    return cfg


@contextlib.contextmanager
def _column_fft():
    """Plans created inside use the column FFT (K5 first / second generation) instead of the column sweep: the
    transposing slab paths run that kernel and are compared bit for bit with the single-GPU step that runs it too.
    (The switch is read when a plan is created; the engine caches plans.)"""
    from jax_ib_b200 import engine
    old = os.environ.get("JAXIB_COL_SWEEP")
    os.environ["JAXIB_COL_SWEEP"] = "0"
    engine._PLANS.clear()
    try:
        yield
    finally:
        if old is None:
            del os.environ["JAXIB_COL_SWEEP"]
        else:
            os.environ["JAXIB_COL_SWEEP"] = old
        engine._PLANS.clear()


def _close(got, ref, tol=5e-6):
    """Sweep-based slabs associate the carried values per slab, one GPU per 32-lane scan: same operator, float32
    round-off apart.  Bars: rel-L2 5e-6 on the velocities, 10x that on the pressure (|p| << |u| in these flows)."""
    for name, a, b in zip("uvp", got, ref):
        assert torch.isfinite(a).all()
        err = float((a - b).norm() / b.norm())
        assert err < (10 * tol if name == "p" else tol), (name, err, float((a - b).abs().max()))


def _single_gpu(cfg, steps):
    from jax_ib_b200 import configs, demo, kinematics
    stp = demo.make_stepper(cfg)
    u, v, p = (torch.from_numpy(a).cuda() for a in configs.initial_fields(cfg))
    table = None
    if cfg["bodies"] is not None:
        tab, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], steps)
        table = torch.from_numpy(np.ascontiguousarray(tab[:, None])).cuda()
    stp.run_block(u, v, p, steps, table)
    return u, v, p


def _virtual(cfg, steps, world, chunks=(None,), peer=False):
    from jax_ib_b200 import configs, demo
    st = slab.SlabStepper(demo.make_spec(cfg), demo.make_particles(cfg), world=world, virtual=True, peer=peer)
    st.load(*(torch.from_numpy(a) for a in configs.initial_fields(cfg)))
    done = 0
    for c in chunks:
        k = steps - done if c is None else c
        st.advance(k)
        done += k
    assert done == steps
    return st.gather()


@pytest.mark.gpu
@pytest.mark.parametrize("n,world", [(256, 1), (256, 2), (256, 4), (1024, 2), (1024, 8)])
def test_virtual_slabs_match_single_gpu_bitwise(n, world):
    cfg = _seam_cfg(n)
    steps = 6
    with _column_fft():
        ref = _single_gpu(cfg, steps)
    got = _virtual(cfg, steps, world, chunks=(2, None))
    for name, a, b in zip("uvp", got, ref):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b), (name, float((a - b).abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("n,world", [(1024, 1), (1024, 2), (1024, 4), (1024, 8), (2048, 4)])
def test_peer_memory_slabs_match_single_gpu_bitwise(n, world, monkeypatch):
    """csrc/slab.cu, transpose path (JAXIB_SLAB_FFT=1): the kernels scatter into the consuming slab's buffers (here:
    all on one GPU, phases in lockstep without the flag barrier).  Same arithmetic per cell as on one GPU with the
    column FFT => identical bits."""
    cfg = _seam_cfg(n)
    steps = 6
    monkeypatch.setenv("JAXIB_SLAB_FFT", "1")
    with _column_fft():
        ref = _single_gpu(cfg, steps)
        got = _virtual(cfg, steps, world, chunks=(2, None), peer=True)
    for name, a, b in zip("uvp", got, ref):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b), (name, float((a - b).abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("n,world", [(1024, 1), (1024, 2), (1024, 4), (1024, 8), (2048, 8)])
def test_peer_memory_slabs_with_the_column_sweep_match_single_gpu(n, world):
    """Default peer-memory path: no transposes; every slab sweeps its own rows of every column and the slabs only
    exchange their carried-value maps (poisson_sweep.cu, passes A / B).  Bodies on the seam, on a slab boundary
    and off the grid as in the bitwise tests."""
    cfg = _seam_cfg(n)
    steps = 6
    ref = _single_gpu(cfg, steps)
    got = _virtual(cfg, steps, world, chunks=(2, None), peer=True)
    _close(got, ref)


@pytest.mark.gpu
def test_peer_memory_slabs_8192_rows_split_column_transform(monkeypatch):
    """Nx = 8192, 520 spectrum columns per slab: the radix 16x32x16 column kernel (one CTA per SM, 260 CTAs)
    runs in two launches of whole waves and the push of the first part overlaps the second launch on a side
    stream (csrc/slab.cu phase 2)."""
    from jax_ib_b200 import configs
    cfg = configs.ib_periodic(8192, 1, dx=0.025)
    cfg["shape"] = (8192, 2048)
    cfg["domain"] = ((0.0, 8192 * 0.025), (0.0, 2048 * 0.025))
    cfg["bodies"]["centers"] = [[0.5 * 8192 * 0.025, 0.5 * 2048 * 0.025]]  # on the slab boundary
    assert slab.SlabLayout(8192, 2048, 2, 0, 16).cpr == 520
    steps = 3
    monkeypatch.setenv("JAXIB_SLAB_FFT", "1")
    with _column_fft():
        ref = _single_gpu(cfg, steps)
        got = _virtual(cfg, steps, 2, peer=True)
    for name, a, b in zip("uvp", got, ref):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
    monkeypatch.delenv("JAXIB_SLAB_FFT")
    _close(_virtual(cfg, steps, 2, peer=True), _single_gpu(cfg, steps))  # and the sweep path on the same case


@pytest.mark.gpu
def test_peer_memory_slabs_without_bodies():
    from jax_ib_b200 import configs
    cfg = dict(configs.ib_periodic(1024, 1), bodies=None)
    ref = _single_gpu(cfg, 4)
    got = _virtual(cfg, 4, 4, peer=True)
    _close(got, ref)


@pytest.mark.gpu
def test_peer_memory_mode_rejects_unsupported_sizes():
    from jax_ib_b200 import _lib, configs
    cfg = configs.small_periodic(64)
    with pytest.raises(_lib.JaxibError):
        _virtual(cfg, 1, 2, peer=True)


@pytest.mark.gpu
def test_virtual_slabs_without_bodies():
    from jax_ib_b200 import configs
    cfg = configs.small_periodic(64, bodies=False)
    with _column_fft():
        ref = _single_gpu(cfg, 5)
    got = _virtual(cfg, 5, 2)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_slabs_on_several_gpus_match_single_gpu(mode):
    """Real exchanges: NCCL collectives, or NVLink stores from the kernels + flag barriers (peer mode).
    Needs >= 2 GPUs on the box, one rank per GPU."""
    import subprocess
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n_gpus >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29100 + os.getpid() % 500), os.path.join(root, "tools", "slab_check.py"),
           "--size", "1024", "--steps", "6"] + (["--peer"] if mode == "peer" else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SLAB_CHECK OK" in res.stdout


# ---- host hints of the peer-memory slab step (pure host logic) --------------------------------------------------
def test_slab_hints_are_conservative_supersets():
    """`jaxib_slab.seam_idle` / `body_first, body_count` (slab.seam_is_idle, slab.bodies_near): every body whose markers
    can reach a slab's rows (radius + the 2 R + 4 halo rows the kernels read) during the block must be inside the
    hinted range; bodies far away are left out; the first / last slab own everything off the grid."""
    dx, R, nx, world = 0.025, 6, 8192, 8
    nloc = nx // world
    rng = np.random.default_rng(1)
    radius = 0.5
    base = (np.arange(64) // 8 + 0.5) * 25.6                       # the G4 lattice: 8 bodies per slab, sorted by x
    xc = base[None, :] + 2.8 * np.sin(np.linspace(0, 1.0, 50))[:, None] * rng.uniform(-1, 1, 64)[None, :]
    assert slab.seam_is_idle(xc, radius, R, dx, 0.0, nx)
    assert not slab.seam_is_idle(np.concatenate([xc, np.full((50, 1), 0.3)], 1), radius, R, dx, 0.0, nx)
    assert not slab.seam_is_idle(np.concatenate([xc, np.full((50, 1), nx * dx - 0.3)], 1), radius, R, dx, 0.0, nx)
    for r in range(world):
        b0, b1 = slab.bodies_near(xc, radius, R, dx, 0.0, r * nloc, (r + 1) * nloc, r == 0, r == world - 1)
        # brute force: rows any marker of body k can touch (centre track +- radius, +- the halo the kernels read)
        lo = (xc.min(0) - radius) / dx - (2 * R + 4)
        hi = (xc.max(0) + radius) / dx + (2 * R + 4)
        must = np.nonzero((hi >= r * nloc) & (lo <= (r + 1) * nloc))[0]
        assert must.size and b0 <= must.min() and must.max() < b1
        assert b1 - b0 <= 24                                       # and it does prune: 8-16 of the 64 bodies (+ neighbours)
    # bodies off the grid belong to the first / last slab; unordered bodies give the covering range
    odd = np.array([[-0.4, 100.0, 205.5, 50.0]])
    assert slab.bodies_near(odd, radius, R, dx, 0.0, 0, nloc, True, False) == (0, 1)
    assert slab.bodies_near(odd, radius, R, dx, 0.0, 7 * nloc, 8 * nloc, False, True) == (2, 3)
    assert slab.bodies_near(odd, radius, R, dx, 0.0, 3 * nloc, 5 * nloc, False, False) == (1, 2)   # 100.0 = row 4000
    assert slab.bodies_near(odd, radius, R, dx, 0.0, 1 * nloc, 2 * nloc, False, False) == (3, 4)   # 50.0 = row 2000
    assert slab.bodies_near(odd[:, :1], radius, R, dx, 0.0, 0, nloc, True, False) is None          # one body: no range
