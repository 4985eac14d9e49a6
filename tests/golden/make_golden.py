"""Generates the golden fixtures in this directory from the ORACLE (oracle/, float32, the
reference's dense O(M*Nx*Ny) Gaussian sums exactly as IBM_Force.py / convolution_functions.py do).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

These are outputs of the restatement.  Since round 2 each of the three reference configurations is tied to a run of
the reference's OWN files / notebooks on the NumPy stand-in for jax (make_reference_notebook_golden.py,
make_reference_workload_golden.py; tests/test_reference_pinned.py compares: flap600 1e-7, bearing300 1e-6,
channel1024x256 bit-identical).  Each file holds a crop
of (u, v, p) around the bodies plus full-field norms / sums, the clock, the body centres and the
step count, so the fixtures stay small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from jax_ib_b200 import configs  # noqa: E402
from _adapter import oracle_state, oracle_step  # noqa: E402

CASES = {
    # name: (config, steps, dense IB?, crop)
    "flap600_20": (configs.flap600(), 20, True, (slice(380, 540), slice(220, 380))),
    "bearing300_20": (configs.bearing300(), 20, True, (slice(70, 230), slice(70, 230))),
    "channel1024x256_5": (configs.channel(), 5, False, (slice(0, 160), slice(0, 256))),
    "small64_50": (configs.small_periodic(64), 50, True, (slice(0, 64), slice(0, 64))),
}


def summarize(st, crop):
    out = {}
    for name, var in (("u", st.velocity[0]), ("v", st.velocity[1]), ("p", st.pressure)):
        d = var.data
        out[name + "_crop"] = d[crop].astype(np.float32)
        out[name + "_l2"] = np.float64(np.sqrt((d.astype(np.float64) ** 2).sum()))
        out[name + "_sum"] = np.float64(d.astype(np.float64).sum())
    out["time"] = np.float32(st.velocity[0].bc.time_stamp)
    out["step_count"] = np.int64(st.Step_count)
    if st.particles is not None:
        out["centers"] = np.asarray(st.particles.centers, dtype=np.float32)
    return out


def main(only=None):
    for name, (cfg, steps, dense, crop) in CASES.items():
        if only and name not in only:
            continue
        st = oracle_state(cfg)
        step = oracle_step(cfg, ib_window=None if dense else 6)
        for _ in range(steps):
            st = step(st)
        data = summarize(st, crop)
        # float32 round-off floor of this trajectory: the same oracle in float64 against float32
        # (what two correct float32 implementations may legitimately differ by)
        st64 = oracle_state(cfg, np.float64)
        for _ in range(steps):
            st64 = step(st64)
        for key, a, b in (("u", st.velocity[0], st64.velocity[0]), ("v", st.velocity[1], st64.velocity[1]),
                          ("p", st.pressure, st64.pressure)):
            d = b.data[crop]
            data[key + "_noise"] = np.float64(np.linalg.norm(a.data[crop] - d) / max(np.linalg.norm(d), 1e-300))
        data["crop"] = np.array([crop[0].start, crop[0].stop, crop[1].start, crop[1].stop])
        data["steps"] = np.int64(steps)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: (v.shape if hasattr(v, "shape") and v.ndim else float(v)) for k, v in data.items()})


if __name__ == "__main__":
    main(sys.argv[1:])
