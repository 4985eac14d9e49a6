"""Runs the REFERENCE'S OWN NOTEBOOKS -- /root/reference/jax_ib/notebooks/Flapping_Demo.ipynb (code cells 1, 3, 5, 8)
and journal_bearing_demo.ipynb (cells 1, 4, 6, 8, 10), executed verbatim (imports; flow conditions, grid, boundary
conditions; the bodies and their kinematics; [tracers;] step function and roll-out) -- on the NumPy stand-in for jax
(oracle/jax_standin), with only the two step counts overridden (inner_steps = 10, outer_steps = 2 instead of
1000 x 10: a 600 x 600 step with 298 markers costs ~4 s on NumPy).  Build container only.

    python tests/golden/make_reference_notebook_golden.py      # ~80 s; writes tests/golden/reference_notebook_{flap600,bearing300}.npz

The notebook is read from the reference tree at run time; nothing of it is copied into this repository.  Stored: the
crop of (u, v, p) that tests/golden/flap600_20.npz uses, full-field norms, body centre, clock, step count, and the
snapshots' step counts -- the same 20-step runs those fixtures hold from the oracle and that the GPU tests replay
(tests/test_dropin_api.py, tests/test_gpu_parity.py).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


NOTEBOOKS = {
    # fixture name: (notebook, code cells executed before the roll-out cell, roll-out cell, the oracle's fixture of the same run)
    "flap600": ("Flapping_Demo.ipynb", (1, 3, 5), 8, "flap600_20.npz"),
    # journal bearing: two concentric circles on the Fourier laws + 1401 passive tracers.  The tracers' pair potential
    # (jax_md.smap.pair + quantity.force) is NOT restated -- they feel no wall repulsion here and their positions are
    # not stored; they are passive, so the FLOW (u, v, p, centres, clock) is exactly what the notebook computes.
    "bearing300": ("journal_bearing_demo.ipynb", (1, 4, 6, 8), 10, "bearing300_20.npz"),
}


def run_notebook(name, inner_steps=10, outer_steps=2):
    sys.dont_write_bytecode = True
    if os.path.join(ROOT, "oracle", "jax_standin") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_standin"))
        sys.path.insert(0, REF)
    import jax.random
    jax.random.ZERO_TEMPERATURE_ONLY = True       # kT = 0 in the notebook: the deviates are multiplied by zero
    jax.random.NUMPY_UNIFORM_FOR_INPUTS = True    # tracer start positions: inputs that do not influence the flow
    notebook, setup_cells, rollout_cell, _ = NOTEBOOKS[name]
    nb = json.load(open(f"{REF}/jax_ib/notebooks/{notebook}"))
    cells = {i: "".join(c["source"]) for i, c in enumerate(nb["cells"]) if c["cell_type"] == "code"}
    ns = {}
    for i in setup_cells:
        exec(compile(cells[i], f"{notebook}[{i}]", "exec"), ns)
    ns["inner_steps"], ns["outer_steps"] = inner_steps, outer_steps  # the only change: 20 steps instead of 10 000
    exec(compile(cells[rollout_cell], f"{notebook}[{rollout_cell}]", "exec"), ns)
    return ns


def main():
    for name, (notebook, _, _, oracle_fixture) in NOTEBOOKS.items():
        t0 = time.time()
        ns = run_notebook(name)
        final, traj = ns["final_result"], ns["trajectory"]
        gold = np.load(os.path.join(HERE, oracle_fixture))
        a, b, c, d = (int(x) for x in gold["crop"])
        out = {"crop": gold["crop"], "steps": np.int64(final.Step_count),
               "time": np.asarray(final.velocity[0].bc.time_stamp, dtype=np.float32),
               "centers": np.asarray(final.particles.particle_center, dtype=np.float32),
               "snapshot_steps": np.asarray(traj.Step_count), "dt": np.float64(ns["dt"]), "shape": np.array(ns["size"])}
        for key, var in (("u", final.velocity[0]), ("v", final.velocity[1]), ("p", final.pressure)):
            full = np.asarray(var.data, dtype=np.float32)
            out[key + "_crop"] = full[a:b, c:d]
            out[key + "_l2"] = np.float64(np.sqrt((full.astype(np.float64) ** 2).sum()))
            out[key + "_sum"] = np.float64(full.astype(np.float64).sum())
        if name == "bearing300":  # passive tracers (no pair force here, see NOTEBOOKS): start, end, mobility mask
            out["md_pos0"] = np.asarray(ns["Pos_all"], dtype=np.float32)
            out["md_pos"] = np.asarray(final.MD_var.position, dtype=np.float32)
            out["md_mobile"] = np.asarray(ns["is_mobile"]).astype(bool)
            out["md_gamma"] = np.float64(ns["gamma_val"])
        path = os.path.join(HERE, f"reference_notebook_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
