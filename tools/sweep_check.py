"""Relative error of the pressure solve against the float64 oracle for one shape (scratch tool; env knobs apply)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _adapter import *  # noqa
from jax_ib_b200 import ops, demo
from oracle import poisson
import test_gpu_parity as T

def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    cfg = T.make_cfg((nx, ny), "periodic", "upwind")
    u, v, p = T.random_fields((nx, ny), 21)
    st64 = T.oracle_state(cfg, np.float64, (u, v, p))
    q64 = poisson.solve_fast_diag(st64.velocity); q64 -= q64.mean()
    plan = ops.Plan(demo.make_spec(cfg))
    q = T.host(plan.pressure_solve(T.dev(u), T.dev(v)))
    err = q - q64
    knobs = {k: os.environ[k] for k in os.environ if k.startswith("JAXIB_")}
    print((nx, ny), knobs, "rel %.3e finite=%s" % (T.rel_l2(q, q64), np.isfinite(q).all()),
          "worst rows", np.argsort(-np.abs(err).max(1))[:6], "row-err max %.2e" % np.abs(err).max(), "q max %.2e" % np.abs(q64).max())
    # error by ky column of the row spectrum
    E = np.abs(np.fft.rfft(err, axis=1)).max(0)
    print("   worst ky", np.argsort(-E)[:8], E[np.argsort(-E)[:3]])
main()
