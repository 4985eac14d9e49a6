"""N > 1 host logic on CPU with the gloo backend (world_size 2): ensemble sharding covers every
simulation exactly once, per-rank kinematics tables equal the single-process ones, and the
max/sum-over-ranks reductions bench.py relies on behave."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jax_ib_b200 import configs, demo, kinematics, parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_sims, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(n_sims, rank, world)
    params = parallel.ensemble_parameters(n_sims)[lo:hi]
    cfg = configs.flap600()
    sums = []
    for prm in params:
        b = dict(cfg["bodies"])
        b.update(prm)
        table, times, centers = kinematics.body_schedule(demo.make_particles(dict(cfg, bodies=b)), 0.0, cfg["dt"], 8)
        sums.append(float(np.abs(table).sum()))
    mine = torch.zeros(n_sims, dtype=torch.float64)
    mine[lo:hi] = torch.tensor(sums, dtype=torch.float64)
    dist.all_reduce(mine)  # disjoint shards: the sum is the concatenation
    counts = torch.zeros(n_sims)
    counts[lo:hi] = 1
    dist.all_reduce(counts)
    mx = parallel.max_over_ranks(float(rank + 1))
    sm = parallel.sum_over_ranks(float(hi - lo))
    parallel.barrier()
    if rank == 0:
        np.savez(os.path.join(out_dir, "r.npz"), table_sums=mine.numpy(), counts=counts.numpy(), mx=mx, sm=sm)
    dist.destroy_process_group()


def test_shard_range_never_drops_items():
    for n in (0, 1, 7, 8, 298, 1024):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1  # the reference's pmap would drop n % world markers instead


def test_ensemble_sharding_world2(tmp_path):
    n_sims, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_sims, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "r.npz")
    np.testing.assert_array_equal(r["counts"], np.ones(n_sims))
    assert float(r["mx"]) == 2.0 and float(r["sm"]) == n_sims
    cfg = configs.flap600()
    want = []
    for prm in parallel.ensemble_parameters(n_sims):
        b = dict(cfg["bodies"])
        b.update(prm)
        table, _, _ = kinematics.body_schedule(demo.make_particles(dict(cfg, bodies=b)), 0.0, cfg["dt"], 8)
        want.append(float(np.abs(table).sum()))
    np.testing.assert_allclose(r["table_sums"], want, rtol=1e-12)


@pytest.mark.gpu
def test_ensemble_stepper_matches_individual_runs():
    cfg = configs.small_periodic(64)
    params = [dict(displacement_param=[[0.4, 0.5]], rotation_param=[[1.5, 0.7, 0.5, 0.3]]),
              dict(displacement_param=[[0.3, 0.7]], rotation_param=[[1.2, 0.5, 0.7, 1.3]]),
              dict(displacement_param=[[0.5, 0.4]], rotation_param=[[1.7, 0.6, 0.4, 2.3]])]
    ens = parallel.EnsembleStepper(cfg, params)
    ens.advance(4)
    ens.advance(3)
    for k, prm in enumerate(params):
        b = dict(cfg["bodies"])
        b.update(prm)
        one = dict(cfg, bodies=b)
        out = demo.make_stepper(one).advance(demo.make_all_variables(one, device="cuda"), 7)
        assert torch.equal(ens.u[k], out.velocity[0].data)
        assert torch.equal(ens.p[k], out.pressure.data)
        np.testing.assert_array_equal(np.asarray(ens.particles[k].particle_center), np.asarray(out.particles.particle_center))
