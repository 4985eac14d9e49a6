"""Prints how far the CUDA path is from the outputs of the reference's own files (tests/golden/reference_pinned.npz),
case by case -- the numbers behind tests/test_gpu_vs_reference.py.  Run on a B200:  python tools/vs_reference_report.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_gpu_vs_reference as T  # noqa: E402
from jax_ib_b200 import advection, funcutils, particle_class as pc, pressure  # noqa: E402

G, host, rel = T.GOLD, T.host, T.rel_l2
out = {}
for tag in ("periodic", "channel"):
    U, V, P = T._variables(G["u"], G["v"], G["p"], *T._bcs(tag))
    for name, fn in (("upwind", advection.advect_upwind), ("van_leer", advection.advect_van_leer),
                     ("van_leer_limiters", advection.advect_van_leer_using_limiters)):
        errs = []
        for k, c in enumerate((U, V)):
            want = G[f"{tag}_{name}_{k}"]
            errs.append(float(np.abs(host(fn(c, (U, V), float(G["dt"])).data) - want).max() / np.abs(want).max()))
        out[f"advect_{name}_{tag}_maxabs_over_scale"] = max(errs)
for tag, solve in (("periodic", pressure.solve_fast_diag), ("channel", pressure.solve_fast_diag_moving_wall),
                   ("far_field", pressure.solve_fast_diag_Far_Field)):
    U, V, P = T._variables(G["u"], G["v"], G["p"], *T._bcs(tag))
    out[f"solve_{tag}_q"] = rel(host(solve((U, V)).data), G[f"{tag}_q"])
    new = pressure.projection_and_update_pressure(pc.All_Variables(None, (U, V), P, [], 0, None), solve)
    out[f"project_{tag}_u"] = rel(host(new.velocity[0].data), G[f"{tag}_proj_u"])
    out[f"project_{tag}_p"] = rel(host(new.pressure.data), G[f"{tag}_proj_p"])
U, V, P = T._variables(G["u"], G["v"], G["p"], *T._bcs("periodic", float(G["ib_t0"])))
fx, fy = T.IBM(pc.All_Variables(T._body(2), (U, V), P, [], 0, None), float(G["dt"]))
out["ib_force_two_bodies_fx"], out["ib_force_two_bodies_fy"] = rel(host(fx.data), G["ib_fx"]), rel(host(fy.data), G["ib_fy"])
for tag in ("periodic", "channel", "far_field"):
    step_fn, av = T._step_case(tag)
    one = step_fn(av)
    six = funcutils.repeated(step_fn, int(G["step_n"]))(av)
    out[f"step1_{tag}"] = {k: rel(host(x), G[f"step1_{tag}_{k}"]) for k, x in
                           (("u", one.velocity[0].data), ("v", one.velocity[1].data), ("p", one.pressure.data))}
    out[f"step6_{tag}"] = {k: rel(host(x), G[f"step_{tag}_{k}"]) for k, x in
                           (("u", six.velocity[0].data), ("v", six.velocity[1].data), ("p", six.pressure.data))}
    out[f"step6_{tag}"]["centre_abs"] = float(np.abs(np.asarray(six.particles.particle_center) - G[f"step_{tag}_centers"]).max())
print(json.dumps(out, indent=1, default=float))
