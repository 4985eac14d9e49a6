"""Timing sweep of rows-per-CTA-group for the fast row kernels K4 / K6 (measurement aid, 1 GPU).

    python tools/tune_rows.py NX NY [NX_GLOBAL]     # NX_GLOBAL > 0: slab mode (open rows), NX = rows of the slab
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    nxg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    from jax_ib_b200 import ops

    dev = torch.device("cuda")
    rows = nx + 2  # room for the open-row halos
    u = torch.randn(rows, ny, device=dev)
    v = torch.randn(rows, ny, device=dev)
    p = torch.randn(rows, ny, device=dev)
    sp = ny // 2 + 4
    spec = torch.zeros(rows, sp, 2, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for which in ("FWD", "INV"):
        res = []
        for rpg in range(1, 13):
            os.environ["JAXIB_RPG_" + which] = str(rpg)
            plan = ops.Plan(ops.GridSpec(nx=nx, ny=ny, step=(0.025, 0.025), dt=5e-4, nu=0.05, nx_global=nxg), 0, dev,
                            alloc_scratch=False)
            ts = []
            for rep in range(12):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if which == "FWD":
                    plan.rows_fwd(u[1:nx + 1], v[1:nx + 1], spec)
                else:
                    plan.rows_inv(spec, u[1:nx + 1], v[1:nx + 1], p[1:nx + 1], u[1:nx + 1], v[1:nx + 1], p[1:nx + 1])
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            res.append((rpg, round(ts[len(ts) // 2], 1)))
            del plan
        os.environ.pop("JAXIB_RPG_" + which)
        print(f"{nx}x{ny} nx_global={nxg} {which}: " + " ".join(f"{r}:{t}" for r, t in res), flush=True)


if __name__ == "__main__":
    main()
