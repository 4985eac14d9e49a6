"""Scratch diagnostic (GPU): Poisson solve error of the CUDA path and of the float32 oracle, both
against a float64 oracle solve, for a list of shapes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from jax_ib_b200 import demo, ops  # noqa: E402
from oracle import poisson  # noqa: E402
from _adapter import oracle_state, rel_l2  # noqa: E402

shapes = [(64, 64), (60, 100), (128, 64), (256, 512), (600, 600), (1024, 256), (2048, 2048)]
for shape in shapes:
    nx, ny = shape
    cfg = dict(name="t", shape=shape, domain=((0.5, 0.5 + 0.05 * nx), (0.25, 0.25 + 0.04 * ny)), dt=1e-3, density=1.3,
               viscosity=0.07, bc="periodic", convect="upwind", force=(0, 0), wall_u=(0, 0), wall_v=(0, 0), bodies=None,
               init="zero")
    rng = np.random.default_rng(21)
    u, v, p = ((rng.standard_normal(shape)).astype(np.float32) for _ in range(3))
    q64 = poisson.solve_fast_diag(oracle_state(cfg, np.float64, (u, v, p)).velocity)
    q64 = q64 - q64.mean()
    q32 = poisson.solve_fast_diag(oracle_state(cfg, np.float32, (u, v, p)).velocity)
    plan = ops.Plan(demo.make_spec(cfg))
    q = plan.pressure_solve(torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()).cpu().numpy()
    print(shape, "gpu-vs-f64 %.2e  f32oracle-vs-f64 %.2e  gpu-vs-f32oracle %.2e  mean(q) %.2e" % (
        rel_l2(q, q64), rel_l2(q32, q64), rel_l2(q, q32), q.mean()))
