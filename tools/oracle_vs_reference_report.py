"""Prints how far the ORACLE is from the outputs of the reference's own files / notebooks, fixture by fixture (CPU; needs
only the committed fixtures under tests/golden/).  python tools/oracle_vs_reference_report.py > profiles/...json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_reference_pinned as T  # noqa: E402
from _adapter import oracle_state, oracle_step, rel_l2  # noqa: E402
from jax_ib_b200 import configs  # noqa: E402
from oracle import field, poisson, stencils, stepper  # noqa: E402

G, F = T.GOLD, np.float32
gold = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name))
out = {}

# stage level (reference_pinned.npz)
for tag in ("periodic", "channel", "far_field"):
    U, V, P = T._velocity(tag)
    dt = float(G["dt"])
    out[f"stencils_{tag}_bit_identical"] = bool(all(
        np.array_equal(fn(c, (U, V), dt), G[f"{tag}_{name}_{k}"])
        for name, fn in (("upwind", stencils.advect_upwind), ("van_leer", stencils.advect_van_leer),
                         ("van_leer_limiters", stencils.advect_van_leer_using_limiters))
        for k, c in enumerate((U, V))) and np.array_equal(stencils.divergence((U, V)), G[f"{tag}_divergence"]))
    q = poisson.SOLVERS[tag]((U, V))
    out[f"solve_{tag}"] = {"rel_l2": rel_l2(q, G[f"{tag}_q"]), "bit_identical": bool(np.array_equal(q, G[f"{tag}_q"]))}

# assembled steps (one flapping body, 1 and 6 steps)
for tag in ("periodic", "channel", "far_field"):
    convect, fns, vals, force = T.STEP_CASES[tag]
    g = T._grid()
    zero = [lambda t: 0.0] * 4
    make = {"periodic": lambda v, f: field.periodic_bc(F(0.0), f), "channel": lambda v, f: field.channel_bc(v, F(0.0), f),
            "far_field": lambda v, f: field.far_field_bc(v, F(0.0), f)}[tag]
    U = field.Var(G["step_u0"], (1.0, 0.5), g, make(vals, fns if fns is not None else zero))
    V = field.Var(G["step_v0"], (0.5, 1.0), g, make(((0.0, 0.0), (0.0, 0.0)), zero))
    P = field.Var(np.zeros(g.shape, F), (0.5, 0.5), g, field.pressure_bc_from_velocity((U, V)))
    forcing = None if force is None else (lambda vv: tuple(F(f) * np.ones_like(x.data) for f, x in zip(force, vv)))
    step = stepper.make_step(float(G["dt"]), float(G["step_viscosity"]), float(G["step_density"]), convect, tag, forcing)
    st = stepper.State(T._bodies(1), (U, V), P)
    for n in range(int(G["step_n"])):
        st = step(st)
        if n == 0:
            out[f"step1_{tag}"] = {k: rel_l2(x, G[f"step1_{tag}_{k}"]) for k, x in
                                   (("u", st.velocity[0].data), ("v", st.velocity[1].data), ("p", st.pressure.data))}
    out[f"step6_{tag}"] = {k: rel_l2(x, G[f"step_{tag}_{k}"]) for k, x in
                           (("u", st.velocity[0].data), ("v", st.velocity[1].data), ("p", st.pressure.data))}
    out[f"step6_{tag}"]["centre_abs"] = float(np.abs(st.particles.centers - G[f"step_{tag}_centers"]).max())

# trajectory fixtures of the oracle against runs of the reference's notebooks / files
for name, mine, theirs in (("flap600_20", "flap600_20.npz", "reference_notebook_flap600.npz"),
                           ("bearing300_20", "bearing300_20.npz", "reference_notebook_bearing300.npz"),
                           ("channel1024x256_5", "channel1024x256_5.npz", "reference_channel1024x256_5.npz"),
                           ("small64_50", "small64_50.npz", "reference_small64_50.npz")):
    a, b = gold(mine), gold(theirs)
    out[f"fixture_{name}"] = {k: rel_l2(a[k + "_crop"], b[k + "_crop"]) for k in "uvp"}
    out[f"fixture_{name}"]["clock_identical"] = bool(F(a["time"]) == F(b["time"]))
    if "centers" in b.files:
        out[f"fixture_{name}"]["centres_identical"] = bool(np.array_equal(a["centers"], b["centers"]))

# the bench workload: windowed oracle against the reference's dense three steps
ref = gold("reference_ib2048_3.npz")
cfg = configs.ib_periodic(2048, 2)
st = oracle_state(cfg)
step = oracle_step(cfg, ib_window=6)
for _ in range(int(ref["steps"])):
    st = step(st)
out["ib2048_3"] = {}
for key, var in (("u", st.velocity[0]), ("v", st.velocity[1]), ("p", st.pressure)):
    crops = np.stack([var.data[a:b, c:d] for a, b, c, d in ref["boxes"]])
    blocks = var.data.astype(np.float64).reshape(64, 32, 64, 32).mean(axis=(1, 3))
    out["ib2048_3"][key] = {"crops": rel_l2(crops, ref[key + "_crops"]), "block_means": rel_l2(blocks, ref[key + "_blocks"])}
out["ib2048_3"]["centre_abs"] = float(np.abs(st.particles.centers - ref["centers"]).max())
print(json.dumps(out, indent=1, default=float))
