"""Per-slab kernel times of the decomposed step on ONE GPU (virtual slabs, stage-level API, no NVLink):
which slab is the slowest in each phase?   python tools/slab_phase_times.py [N] [WORLD]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    from jax_ib_b200 import configs, demo, kinematics, ops, slab

    cfg = configs.ib_periodic(n, max(1, n // 2048))
    st = slab.SlabStepper(demo.make_spec(cfg), demo.make_particles(cfg), world=world, virtual=True)
    st.load(*(torch.from_numpy(a) for a in configs.initial_fields(cfg)))
    st.advance(2)
    table, _ = st.schedule(1)
    state = table[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=5):
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return round(sorted(ts)[len(ts) // 2], 1)

    for r in st.ranks:
        L = r.lay
        k1 = timed(lambda: ops.explicit_predict(r.spec_ext, r.u, r.v, r.p, out=(r.us, r.vs)))
        k2 = timed(lambda: ops.ib_marker_force(r.spec_ib, r.bodies, state, r._rows(r.us, r.seg), r._rows(r.vs, r.seg), r.markers))
        k2t = timed(lambda: ops.ib_marker_force(r.spec_top, r.bodies, state, r._rows(r.us, r.top), r._rows(r.vs, r.top), r.markers)) if r.top is not None else 0.0
        k3 = timed(lambda: ops.ib_spread_bodies(r.spec_ib, r.bodies, state, r.markers, r._rows(r.us, r.seg), r._rows(r.vs, r.seg)))
        k3s = timed(lambda: ops.ib_spread_bodies(r.spec_seam, r.bodies, state, r.markers, r._rows(r.us, r.seam), r._rows(r.vs, r.seam))) if r.seam is not None else 0.0
        k4 = timed(lambda: r.plan_rows.rows_fwd(r.owned(r.us), r.owned(r.vs), r.spectrum, sp=L.sp))
        k5 = timed(lambda: r.plan_cols.cols(r.recv, sp=L.cpr, ncols=L.cpr, col0=L.rank * L.cpr))
        po = r.owned(r.p)
        k6 = timed(lambda: r.plan_rows.rows_inv(r.spectrum, r.owned(r.us), r.owned(r.vs), po, r.owned(r.u), r.owned(r.v), po, sp=L.sp))
        print(f"slab {L.rank}: K1 {k1}  K2 {k2} (+top {k2t})  K3 {k3} (+seam {k3s})  K4 {k4}  K5 {k5}  K6 {k6}  "
              f"sum {round(k1 + k2 + k2t + k3 + k3s + k4 + k5 + k6, 1)} us", flush=True)


if __name__ == "__main__":
    main()
