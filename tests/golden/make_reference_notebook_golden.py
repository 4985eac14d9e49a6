"""Runs the REFERENCE'S OWN NOTEBOOK -- /root/reference/jax_ib/notebooks/Flapping_Demo.ipynb, code cells 1, 3, 5 and 8
executed verbatim (imports; flow conditions, grid, boundary conditions; the ellipse and its kinematics; step function
and rollout) -- on the NumPy stand-in for jax (oracle/jax_standin), with only the two step counts overridden
(inner_steps = 10, outer_steps = 2 instead of 1000 x 10: each step costs ~4 s on NumPy).  Build container only.

    python tests/golden/make_reference_notebook_golden.py      # ~90 s; writes tests/golden/reference_notebook_flap600.npz

The notebook is read from the reference tree at run time; nothing of it is copied into this repository.  Stored: the
crop of (u, v, p) that tests/golden/flap600_20.npz uses, full-field norms, body centre, clock, step count, and the
snapshots' step counts -- the same 20-step run that fixture holds from the oracle and that the GPU tests replay through
the reference-shaped API (tests/test_dropin_api.py::test_flapping_demo_through_the_reference_api).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def run_notebook(inner_steps=10, outer_steps=2):
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_standin"))
    sys.path.insert(0, REF)
    nb = json.load(open(REF + "/jax_ib/notebooks/Flapping_Demo.ipynb"))
    cells = {i: "".join(c["source"]) for i, c in enumerate(nb["cells"]) if c["cell_type"] == "code"}
    ns = {}
    for i in (1, 3, 5):
        exec(compile(cells[i], f"Flapping_Demo.ipynb[{i}]", "exec"), ns)
    ns["inner_steps"], ns["outer_steps"] = inner_steps, outer_steps  # the only change: 20 steps instead of 10 000
    exec(compile(cells[8], "Flapping_Demo.ipynb[8]", "exec"), ns)
    return ns


def main():
    t0 = time.time()
    ns = run_notebook()
    final, traj = ns["final_result"], ns["trajectory"]
    gold = np.load(os.path.join(HERE, "flap600_20.npz"))
    a, b, c, d = (int(x) for x in gold["crop"])
    out = {"crop": gold["crop"], "steps": np.int64(final.Step_count), "time": np.asarray(final.velocity[0].bc.time_stamp, dtype=np.float32),
           "centers": np.asarray(final.particles.particle_center, dtype=np.float32),
           "snapshot_steps": np.asarray(traj.Step_count), "dt": np.float64(ns["dt"]), "shape": np.array(ns["size"])}
    for key, var in (("u", final.velocity[0]), ("v", final.velocity[1]), ("p", final.pressure)):
        full = np.asarray(var.data, dtype=np.float32)
        out[key + "_crop"] = full[a:b, c:d]
        out[key + "_l2"] = np.float64(np.sqrt((full.astype(np.float64) ** 2).sum()))
        out[key + "_sum"] = np.float64(full.astype(np.float64).sum())
    path = os.path.join(HERE, "reference_notebook_flap600.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
