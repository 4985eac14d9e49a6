"""Prescribed rigid-body kinematics (host side), same names and parameter layouts as the reference's
jax_ib/base/kinematics.py:5-105, written against torch so that `torch.func.jacrev` plays the role of
`jax.jacrev` (IBM_Force.py:101-104, particle_motion.py:47-52).  These run on the HOST in float32:
the per-step body state they produce is a tiny table consumed by the CUDA kernels (`body_schedule`).
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np
import torch

_TWO_PI = 2 * math.pi


def _col(parameters, i):
    return [row[i] for row in parameters]


def _as(x, t):
    return torch.as_tensor(np.asarray(x, dtype=np.float32)) if not isinstance(x, torch.Tensor) else x.to(torch.float32)


def displacement(parameters, t):
    """kinematics.py:5-7."""
    A0, f = (_as(x, t) for x in list(*parameters))
    return torch.stack([A0 / 2 * torch.cos(_TWO_PI * f * t), torch.zeros_like(t)])


def rotation(parameters, t):
    """kinematics.py:10-12."""
    alpha0, beta, f, phi = (_as(x, t) for x in list(*parameters))
    return alpha0 + beta * torch.sin(_TWO_PI * f * t + phi)


def _fourier_parts(parameters, t, n_fields=6, p_index=5):
    alpha0 = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 0)])
    f = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 1)])
    phi = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 2)])
    alpha = torch.stack([_as(x, t).reshape(-1) for x in _col(parameters, 3)])
    beta = torch.stack([_as(x, t).reshape(-1) for x in _col(parameters, 4)])
    p = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, p_index)])
    nb, nt = alpha.shape
    freq = torch.arange(1, nt + 1, dtype=torch.float32).repeat(nb, 1)
    return alpha0, f.reshape(nb, 1), phi.reshape(nb, 1), alpha, beta, p, freq


def Displacement_Foil_Fourier_Dotted_Mutliple(parameters, t):
    """kinematics.py:14-38: returns [2, B]."""
    alpha0, f, phi, alpha, beta, p, freq = _fourier_parts(parameters, t)
    inside = _TWO_PI * t * freq * f + phi
    a1 = (alpha * torch.sin(inside)).sum(dim=1)
    a1 = a1 + p * (beta * torch.cos(inside)).sum(dim=1)
    return torch.stack([-alpha0 * t, a1])


def rotation_Foil_Fourier_Dotted_Mutliple(parameters, t):
    """kinematics.py:41-65: returns [B]."""
    alpha0, f, phi, alpha, beta, p, freq = _fourier_parts(parameters, t)
    inside = _TWO_PI * t * freq * f + phi
    a1 = (alpha * torch.sin(inside)).sum(dim=1)
    a1 = a1 + p * (beta * torch.cos(inside)).sum(dim=1)
    return alpha0 * t + a1


def rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED(parameters, t):
    """kinematics.py:67-105: returns [B]."""
    alpha0, f, phi, alpha, beta, p, freq = _fourier_parts(parameters, t, 7, 6)
    theta_av = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 5)])
    inside = _TWO_PI * t * freq * f + phi
    a1 = (alpha * torch.sin(inside)).sum(dim=1) + p * (beta * torch.cos(inside)).sum(dim=1)
    in2 = _TWO_PI * freq + phi
    a2 = (alpha * torch.sin(in2)).sum(dim=1) + p * (beta * torch.cos(in2)).sum(dim=1)
    in3 = _TWO_PI * freq * 0.0 + phi
    a3 = (alpha * torch.sin(in3)).sum(dim=1) + p * (beta * torch.cos(in3)).sum(dim=1)
    return theta_av * (alpha0 * t + a1) / (alpha0 + a2 - a3)


# ---------------------------------------------------------------------------------------------------
def _values_and_rates(fn: Callable, params, times: torch.Tensor):
    """fn(params, t) and d/dt fn(params, t) for a vector of times (vmap over jacrev, like jax)."""
    f = lambda t: fn(params, t)
    vals = torch.func.vmap(f)(times)
    rates = torch.func.vmap(torch.func.jacrev(f))(times)
    return vals, rates


def body_schedule(particles, t0: float, dt: float, nsteps: int):
    """Per-step rigid-body state for `nsteps` steps starting at simulation time t0.

    Follows IBM_Force.py:27-47,99-104 (state at t_n, per-body parameter evaluation), boundaries.py:530
    (t_{n+1} = t_n + dt in float32) and particle_motion.py:47-55 (centre += dt * Xdot(t_{n+1})).
    Returns (table float32 [nsteps, B, 8], times float32 [nsteps + 1], centres float32 [nsteps + 1, B, 2]).
    """
    centers0 = np.asarray(particles.particle_center, dtype=np.float32).reshape(-1, 2)
    B = centers0.shape[0]
    f32 = np.float32
    times = np.add.accumulate(np.concatenate([[f32(t0)], np.full(nsteps, f32(dt), dtype=f32)]).astype(f32), dtype=f32)
    tt = torch.from_numpy(times.copy())
    disp, rot = particles.Displacement_EQ, particles.Rotation_EQ
    dparams, rparams = particles.displacement_param, particles.rotation_param
    # all-body translation velocity at t_1..t_n drives the centres (particle_motion.py:47-55)
    _, U_all = _values_and_rates(disp, list(dparams), tt)
    U_all = U_all.reshape(nsteps + 1, 2, -1).numpy().astype(f32)
    if U_all.shape[2] == 1 and B > 1:
        U_all = np.repeat(U_all, B, axis=2)
    centers = np.empty((nsteps + 1, B, 2), dtype=f32)
    centers[0] = centers0
    for a in range(2):
        inc = f32(dt) * U_all[1:, a, :]  # float32 product, then sequential float32 accumulation
        centers[:, :, a] = np.add.accumulate(np.concatenate([centers0[None, :, a], inc], axis=0), axis=0, dtype=f32)
    table = np.zeros((nsteps, B, 8), dtype=f32)
    table[:, :, 2:4] = centers[:nsteps]
    # The reference evaluates the laws body by body (`[param_b]`, IBM_Force.py:99-104).  The multi-body
    # laws are element-wise over bodies, so one vectorised evaluation of all bodies gives the same
    # numbers at a fraction of the host cost; single-body laws (`displacement`, `rotation`) take the
    # per-body path.
    th_all = om_all = None
    if B > 1:
        try:
            th_v, om_v = _values_and_rates(rot, list(rparams), tt[:nsteps])
            if th_v.reshape(nsteps, -1).shape[1] == B and U_all.shape[2] == B:
                th_all = th_v.reshape(nsteps, B).numpy().astype(f32)
                om_all = om_v.reshape(nsteps, B).numpy().astype(f32)
        except Exception:
            th_all = None
    if th_all is not None:
        table[:, :, 0] = np.cos(th_all)
        table[:, :, 1] = np.sin(th_all)
        table[:, :, 4:6] = np.transpose(U_all[:nsteps], (0, 2, 1))
        table[:, :, 6] = om_all
        return table, times, centers
    for b in range(B):
        if B == 1:
            U_b = U_all[:nsteps, :, 0]
        else:
            _, U_b = _values_and_rates(disp, [dparams[b]], tt[:nsteps])
            U_b = U_b.reshape(nsteps, 2, -1).numpy().astype(f32)[:, :, 0]
        th_b, om_b = _values_and_rates(rot, [rparams[b]], tt[:nsteps])
        th = th_b.reshape(nsteps, -1).numpy().astype(f32)[:, 0]
        om = om_b.reshape(nsteps, -1).numpy().astype(f32)[:, 0]
        table[:, b, 0] = np.cos(th)
        table[:, b, 1] = np.sin(th)
        table[:, b, 4:6] = U_b
        table[:, b, 6] = om
    return table, times, centers


# ---- multi-body harmonic laws (extension: the reference's `displacement` / `rotation` accept one
# body only -- `list(*parameters)` -- so grids with several flapping bodies need a vectorised form)
def displacement_multiple(parameters, t):
    """[2, B]: per body [A0/2 cos(2 pi f t), 0] with parameters = [[A0, f], ...]."""
    A0 = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 0)])
    f = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 1)])
    x = A0 / 2 * torch.cos(_TWO_PI * f * t)
    return torch.stack([x, torch.zeros_like(x)])


def rotation_multiple(parameters, t):
    """[B]: alpha0 + beta sin(2 pi f t + phi) with parameters = [[alpha0, beta, f, phi], ...]."""
    a0, beta, f, phi = (torch.stack([_as(x, t).reshape(()) for x in _col(parameters, i)]) for i in range(4))
    return a0 + beta * torch.sin(_TWO_PI * f * t + phi)
