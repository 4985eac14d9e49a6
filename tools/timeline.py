"""Kernel timeline of the fused step from CUPTI (torch.profiler): when every kernel of a step starts and ends.

    python tools/timeline.py [workload] [steps]

The step's kernels are chained by programmatic dependent launch, so event brackets cannot see inside a step; the
profiler's activity records can.  A dependent kernel "starts" when its first CTA becomes resident (it may then sit in
griddepcontrol.wait), so the END times are what order the critical path.  Prints, per kernel, the median start / end
in us relative to the predictor's start and the gain of each end over the previous kernel's end.  Diagnostic only:
nothing printed here is a bench value.
"""
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import workload_cfg  # noqa: E402
from jax_ib_b200 import configs, demo, kinematics  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ib2048"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    cfg = workload_cfg(name)
    st = demo.make_stepper(cfg)
    u, v, p = (torch.from_numpy(a).cuda() for a in configs.initial_fields(cfg))
    td = th = None
    if cfg["bodies"] is not None:
        table, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], 10 + K)
        th = np.ascontiguousarray(table[:, None])
        td = torch.from_numpy(th).cuda()
    sl = lambda a, b: (None if td is None else td[a:b], None if th is None else th[a:b])
    d, h = sl(0, 10)
    st.run_block(u, v, p, 10, d, table_host=h)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        d, h = sl(10, 10 + K)
        st.run_block(u, v, p, K, d, table_host=h)
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memset" not in e.name
                  and "Memcpy" not in e.name), key=lambda e: e.time_range.start)
    names = []
    for e in evs:
        if e.name in names:
            break
        names.append(e.name)
    per = len(names)
    nsteps = len(evs) // per
    rows = {n: ([], []) for n in names}
    totals = []
    for s in range(2, nsteps):  # skip the first two steps of the block
        blk = evs[s * per:(s + 1) * per]
        if [e.name for e in blk] != names:
            continue
        t0 = blk[0].time_range.start
        for e in blk:
            rows[e.name][0].append(e.time_range.start - t0)
            rows[e.name][1].append(e.time_range.end - t0)
        if s + 1 < nsteps:
            totals.append(evs[(s + 1) * per].time_range.start - t0)
    knobs = {k: os.environ[k] for k in os.environ if k.startswith("JAXIB_")}
    print(f"timeline {name} {knobs}: {per} kernels per step, {len(totals)} steps; us from the first kernel's start")
    prev_end = 0.0
    for n in names:
        s_, e_ = statistics.median(rows[n][0]), statistics.median(rows[n][1])
        short = n.split("(")[0].split("::")[-1][:44]
        print(f"  {short:44s} start {s_:7.1f}  end {e_:7.1f}  (+{e_ - prev_end:5.1f} after the previous end)")
        prev_end = e_
    if totals:
        print(f"  next step's first kernel starts at {statistics.median(totals):.1f}")


if __name__ == "__main__":
    main()
