"""Prints the sha1 of (u, v, p) after `steps` fused steps of a workload (A/B runs of environment switches).

    python tools/field_checksum.py ib1024 6
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import workload_cfg  # noqa: E402
from jax_ib_b200 import configs, demo, kinematics  # noqa: E402


def checksum(name, steps, with_host_table=True):
    cfg = workload_cfg(name)
    st = demo.make_stepper(cfg)
    u, v, p = (torch.from_numpy(a).cuda() for a in configs.initial_fields(cfg))
    table, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], steps)
    th = np.ascontiguousarray(table[:, None])
    st.run_block(u, v, p, steps, torch.from_numpy(th).cuda(), table_host=th if with_host_table else None)
    torch.cuda.synchronize()
    h = hashlib.sha1()
    for t in (u, v, p):
        h.update(t.cpu().numpy().tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    print("checksum", checksum(sys.argv[1], int(sys.argv[2])))
