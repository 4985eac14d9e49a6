"""Test-side glue: builds ORACLE state from a jax_ib_b200.configs description (the product never
imports oracle/; only tests do)."""
import numpy as np

from jax_ib_b200 import configs
from oracle import field, ib, kinematics as okin, stepper


def oracle_bodies(cfg, dtype=np.float32):
    b = cfg["bodies"]
    if b is None:
        return None
    kw = {"ntheta": b["ntheta"]} if "ntheta" in b else {}
    shape = {"ellipse": ib.ellipse_shape, "circle": ib.circle_shape}[b["shape"]]
    return ib.Bodies(np.asarray(b["centers"], dtype=dtype), b["geometry"], b["displacement_param"], b["rotation_param"],
                     lambda g: shape(g, dtype=dtype, **kw), okin.DISPLACEMENT[b["displacement"]], okin.ROTATION[b["rotation"]])


def oracle_state(cfg, dtype=np.float32, fields=None):
    g = field.Grid(cfg["shape"], cfg["domain"])
    t0 = np.dtype(dtype).type(0.0)
    if cfg["bc"] == "periodic":
        bcu, bcv = field.periodic_bc(t0), field.periodic_bc(t0)
    else:
        bcu = field.channel_bc(((0.0, 0.0), tuple(cfg["wall_u"])), t0)
        bcv = field.channel_bc(((0.0, 0.0), tuple(cfg["wall_v"])), t0)
    u, v, p = fields if fields is not None else configs.initial_fields(cfg)
    U = field.Var(u.astype(dtype), (1.0, 0.5), g, bcu)
    V = field.Var(v.astype(dtype), (0.5, 1.0), g, bcv)
    P = field.Var(p.astype(dtype), (0.5, 0.5), g, field.pressure_bc_from_velocity((U, V)))
    return stepper.State(oracle_bodies(cfg, dtype), (U, V), P)


def oracle_step(cfg, ib_window=6, **kw):
    forcing = None
    if any(cfg["force"]):
        fx, fy = cfg["force"]
        forcing = lambda v: tuple(v_i.data.dtype.type(f) * np.ones_like(v_i.data) for f, v_i in zip((fx, fy), v))
    solve = "periodic" if cfg["bc"] == "periodic" else "channel"
    return stepper.make_step(cfg["dt"], cfg["viscosity"], cfg["density"], cfg["convect"], solve, forcing,
                             ib_window=ib_window, **kw)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (d if d > 0 else 1.0))
