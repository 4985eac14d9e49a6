"""Per-stage CUDA-event timings of the fused step (scratch tool for kernel tuning).

    python tools/stage_times.py [workload] [steps]     # env knobs: JAXIB_ROWS_PER_GROUP, JAXIB_COL_PAIRS, ...
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import workload_cfg  # noqa: E402
from jax_ib_b200 import configs, demo, kinematics, ops  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ib2048"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    flush_on = os.environ.get("FLUSH", "1") == "1"
    cfg = workload_cfg(name)
    st = demo.make_stepper(cfg)
    u, v, p = (torch.from_numpy(a).cuda() for a in configs.initial_fields(cfg))
    table = None
    if cfg["bodies"] is not None:
        table, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], 10 + K)
        table = torch.from_numpy(np.ascontiguousarray(table[:, None])).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush_on else None
    st.run_block(u, v, p, 10, None if table is None else table[:10])
    prof = ops.step_profile(st.plan, u, v, p, K, st.bodies, None if table is None else table[10:], flush=flush,
                            tile_flags=st.tile_flags)
    cells = cfg["shape"][0] * cfg["shape"][1]
    us = {k: round(val / K * 1e3, 2) for k, val in prof.items()}
    knobs = {k: os.environ[k] for k in os.environ if k.startswith("JAXIB_")}
    print(name, knobs, us, "Mcell-steps/s=%.0f" % (cells * K / (prof["step"] * 1e-3) / 1e6),
          "frac=%.3f" % (68.0 * cells * K / (prof["step"] * 1e-3) / 1e9 / 6533.2), flush_on and "flushed" or "warm")


if __name__ == "__main__":
    main()
