"""Whole-step CUDA-event timings of the fused step as bench.py takes them (scratch tool for tuning).

    python tools/step_time.py [workload] [steps]        # env knobs: JAXIB_IB_OVERLAP, JAXIB_K1_W, JAXIB_ROWS, ...
Prints us per step with the L2 flushed before every step (one event pair per step) and back to back.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import workload_cfg  # noqa: E402
from jax_ib_b200 import configs, demo, kinematics  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ib2048"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    cfg = workload_cfg(name)
    st = demo.make_stepper(cfg)
    u, v, p = (torch.from_numpy(a).cuda() for a in configs.initial_fields(cfg))
    th = td = None
    if cfg["bodies"] is not None:
        table, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], 10 + 2 * K)
        th = np.ascontiguousarray(table[:, None])
        td = torch.from_numpy(th).cuda()
    sl = lambda a, b: (None if td is None else td[a:b], None if th is None else th[a:b])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # FLUSH_MODE=clean: after the 256 MiB fill, READ another 256 MiB buffer: the L2 is then cold AND clean, so the step
    # does not also pay for writing the flush buffer's dirty lines back to DRAM (an artefact of flushing by writing)
    clean = os.environ.get("FLUSH_MODE") == "clean"
    flush2 = torch.zeros(32 << 20, dtype=torch.int64, device="cuda") if clean else None
    d, h = sl(0, 10)
    st.run_block(u, v, p, 10, d, table_host=h)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for n in range(K):
        flush.fill_(n & 0xff)
        if clean:
            flush2.sum()
        d, h = sl(10 + n, 11 + n)
        evs[n][0].record()
        st.run_block(u, v, p, 1, d, table_host=h)
        evs[n][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d, h = sl(10 + K, 10 + 2 * K)
    e0.record()
    st.run_block(u, v, p, K, d, table_host=h)
    e1.record()
    torch.cuda.synchronize()
    cells = cfg["shape"][0] * cfg["shape"][1]
    mean = sum(ts) / K
    knobs = {k: os.environ[k] for k in os.environ if k.startswith("JAXIB_") or k == "FLUSH_MODE"}
    print(name, knobs, "flushed us/step mean=%.2f median=%.2f frac=%.3f | back-to-back %.2f frac=%.3f | finite=%s" % (
        mean, ts[K // 2], 68.0 * cells / (mean * 1e-6) / 1e9 / 6533.2, e0.elapsed_time(e1) * 1e3 / K,
        68.0 * cells / (e0.elapsed_time(e1) * 1e-3 / K) / 1e9 / 6533.2, bool(torch.isfinite(u).all())))


if __name__ == "__main__":
    main()
