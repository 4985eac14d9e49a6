"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for tests).

Two ways the path shards:
  * ensemble sweep -- independent simulations are split across ranks with NO data-path collective
    (this module: `shard_range`, `ensemble_parameters`, `EnsembleStepper`);
  * slab decomposition of one large grid -- halo rows + FFT-transpose all-to-all (`slab.py`).

Unlike the reference's marker pmap, which silently drops `M mod device_count` markers
(IBM_Force.py:77-78), `shard_range` never drops items: the remainder goes to the first ranks.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: Optional[str] = None, device: Optional[torch.device] = None):
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if backend == "nccl" and device is not None else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(x: float, device=None) -> float:
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device=None) -> float:
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def ensemble_parameters(n_sims: int, seed: int = 1234):
    """G5 `ens1024`: per-simulation flapping kinematics A0~U[2.0,3.2], f~U[0.15,0.35],
    beta~U[pi/8,3pi/8], phi~U[0,2pi).  Deterministic in (n_sims, seed) and independent of the sharding."""
    rng = np.random.default_rng(seed)
    A0 = rng.uniform(2.0, 3.2, n_sims)
    f = rng.uniform(0.15, 0.35, n_sims)
    beta = rng.uniform(math.pi / 8, 3 * math.pi / 8, n_sims)
    phi = rng.uniform(0.0, 2 * math.pi, n_sims)
    return [dict(displacement_param=[[float(A0[i]), float(f[i])]],
                 rotation_param=[[math.pi / 2, float(beta[i]), float(f[i]), float(phi[i])]]) for i in range(n_sims)]


class EnsembleStepper:
    """`len(params)` independent simulations of one config on ONE GPU, advanced together: every kernel
    launch carries the batch on its leading axis and the kinematics table differs per simulation."""

    def __init__(self, cfg, params: Sequence[dict], device="cuda", ib_window=6):
        from . import demo, engine

        self.cfg, self.params = cfg, list(params)
        self.batch = len(self.params)
        self.device = torch.device(device)
        self.particles = []
        for prm in self.params:
            b = dict(cfg["bodies"])
            b.update(prm)
            self.particles.append(demo.make_particles(dict(cfg, bodies=b)))
        spec = demo.make_spec(cfg, batch=self.batch, ib_window=ib_window)
        self.stepper = engine.FusedStepper(spec, self.particles[0], device=self.device)
        from . import configs

        u, v, p = configs.initial_fields(cfg)
        rep = lambda a: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (self.batch,) + a.shape))).to(self.device)
        self.u, self.v, self.p = rep(u), rep(v), rep(p)
        self.time = np.float32(0.0)
        self.steps_done = 0

    def schedule(self, nsteps: int) -> torch.Tensor:
        from . import kinematics

        tabs, last_centers, times = [], [], None
        for prt in self.particles:
            table, times, centers = kinematics.body_schedule(prt, self.time, self.cfg["dt"], nsteps)
            tabs.append(table)
            last_centers.append(centers[-1])
        self._pending = (times[-1], last_centers)
        return torch.from_numpy(np.ascontiguousarray(np.stack(tabs, axis=1))).to(self.device)

    def advance(self, nsteps: int, table: Optional[torch.Tensor] = None):
        table = self.schedule(nsteps) if table is None else table
        self.stepper.run_block(self.u, self.v, self.p, nsteps, table)
        t_end, centers = self._pending
        self.time = t_end
        self.particles = [dataclasses.replace(p, particle_center=c) for p, c in zip(self.particles, centers)]
        self.steps_done += nsteps
