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


def _time(t):
    """The laws take the clock as the reference does: a Python / NumPy scalar or a tensor (under jacrev)."""
    return t if isinstance(t, torch.Tensor) else torch.as_tensor(np.float32(t))


def displacement(parameters, t):
    """kinematics.py:5-7."""
    t = _time(t)
    A0, f = (_as(x, t) for x in list(*parameters))
    return torch.stack([A0 / 2 * torch.cos(_TWO_PI * f * t), torch.zeros_like(t)])


def rotation(parameters, t):
    """kinematics.py:10-12."""
    t = _time(t)
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
    t = _time(t)
    alpha0, f, phi, alpha, beta, p, freq = _fourier_parts(parameters, t)
    inside = _TWO_PI * t * freq * f + phi
    a1 = (alpha * torch.sin(inside)).sum(dim=1)
    a1 = a1 + p * (beta * torch.cos(inside)).sum(dim=1)
    return torch.stack([-alpha0 * t, a1])


def rotation_Foil_Fourier_Dotted_Mutliple(parameters, t):
    """kinematics.py:41-65: returns [B]."""
    t = _time(t)
    alpha0, f, phi, alpha, beta, p, freq = _fourier_parts(parameters, t)
    inside = _TWO_PI * t * freq * f + phi
    a1 = (alpha * torch.sin(inside)).sum(dim=1)
    a1 = a1 + p * (beta * torch.cos(inside)).sum(dim=1)
    return alpha0 * t + a1


def rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED(parameters, t):
    """kinematics.py:67-105: returns [B]."""
    t = _time(t)
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
# Closed forms of the library's own laws and of their time derivatives, vectorised over the time stamps of a
# block (NumPy float32).  The reference obtains the rates with jax.jacrev at every step (IBM_Force.py:101-104,
# particle_motion.py:47-52); torch.func.vmap(jacrev) reproduces that for arbitrary user laws but costs ~2 ms
# of host time per call, which dominated the end-to-end step.  Same formulas, same float32
# evaluation order as the functions above; user-supplied laws keep the autodiff path.
_F32 = np.float32


def _cols_np(parameters, n):
    cols = list(zip(*parameters))
    return [np.asarray([np.asarray(x, dtype=_F32).reshape(-1) for x in cols[i]], dtype=_F32) for i in range(n)]


def _cf_displacement_multi(parameters, tt):
    A0, f = (c[:, 0] for c in _cols_np(parameters, 2))
    w = _F32(_TWO_PI) * f
    ph = w[None, :] * tt[:, None]
    x = A0[None, :] / _F32(2) * np.cos(ph)
    dx = -(A0[None, :] / _F32(2)) * np.sin(ph) * w[None, :]
    z = np.zeros_like(x)
    return np.stack([x, z], axis=1), np.stack([dx, z], axis=1)  # [n, 2, B]


def _cf_rotation_multi(parameters, tt):
    a0, beta, f, phi = (c[:, 0] for c in _cols_np(parameters, 4))
    w = _F32(_TWO_PI) * f
    ph = w[None, :] * tt[:, None] + phi[None, :]
    return a0[None, :] + beta[None, :] * np.sin(ph), beta[None, :] * np.cos(ph) * w[None, :]  # [n, B]


def _cf_series(parameters, tt, n_fields=6, p_index=5):
    cols = _cols_np(parameters, n_fields)
    alpha0, f, phi = cols[0][:, 0], cols[1][:, 0], cols[2][:, 0]
    alpha, beta, p = cols[3], cols[4], cols[p_index][:, 0]
    nb, nt = alpha.shape
    freq = np.tile(np.arange(1, nt + 1, dtype=_F32), (nb, 1))
    w = _F32(_TWO_PI) * freq * f[:, None]                                        # [B, nt]
    inside = _F32(_TWO_PI) * tt[:, None, None] * freq[None] * f[None, :, None] + phi[None, :, None]  # [n, B, nt]
    a1 = (alpha[None] * np.sin(inside)).sum(axis=2) + p[None] * (beta[None] * np.cos(inside)).sum(axis=2)
    d1 = (alpha[None] * np.cos(inside) * w[None]).sum(axis=2) - p[None] * (beta[None] * np.sin(inside) * w[None]).sum(axis=2)
    return alpha0, a1.astype(_F32), d1.astype(_F32), cols, freq


def _cf_displacement_fourier(parameters, tt):
    alpha0, a1, d1, _, _ = _cf_series(parameters, tt)
    x = -alpha0[None, :] * tt[:, None]
    return np.stack([x, a1], axis=1), np.stack([np.broadcast_to(-alpha0[None, :], a1.shape), d1], axis=1)


def _cf_rotation_fourier(parameters, tt):
    alpha0, a1, d1, _, _ = _cf_series(parameters, tt)
    return alpha0[None, :] * tt[:, None] + a1, alpha0[None, :] + d1


def _cf_rotation_fourier_normalized(parameters, tt):
    alpha0, a1, d1, cols, freq = _cf_series(parameters, tt, 7, 6)
    theta_av, phi, alpha, beta, p = cols[5][:, 0], cols[2][:, 0], cols[3], cols[4], cols[6][:, 0]
    in2 = _F32(_TWO_PI) * freq + phi[:, None]
    a2 = (alpha * np.sin(in2)).sum(axis=1) + p * (beta * np.cos(in2)).sum(axis=1)
    in3 = _F32(_TWO_PI) * freq * _F32(0) + phi[:, None]
    a3 = (alpha * np.sin(in3)).sum(axis=1) + p * (beta * np.cos(in3)).sum(axis=1)
    den = (alpha0 + a2 - a3)[None, :]
    return theta_av[None, :] * (alpha0[None, :] * tt[:, None] + a1) / den, theta_av[None, :] * (alpha0[None, :] + d1) / den


def _closed_form(fn):
    return _CLOSED_FORM.get(getattr(fn, "__name__", None)) if getattr(fn, "__module__", None) == __name__ else None


def _values_and_rates(fn: Callable, params, times: torch.Tensor):
    """fn(params, t) and d/dt fn(params, t) for a vector of times: the closed form for this module's own laws,
    vmap over jacrev (like jax) for anything else."""
    cf = _closed_form(fn)
    if cf is not None:
        try:
            vals, rates = cf(params, np.asarray(times, dtype=_F32))
            return torch.from_numpy(np.ascontiguousarray(vals, dtype=_F32)), torch.from_numpy(np.ascontiguousarray(rates, dtype=_F32))
        except Exception:  # unusual parameter layout: fall through to autodiff
            pass
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
    t = _time(t)
    A0 = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 0)])
    f = torch.stack([_as(x, t).reshape(()) for x in _col(parameters, 1)])
    x = A0 / 2 * torch.cos(_TWO_PI * f * t)
    return torch.stack([x, torch.zeros_like(x)])


def rotation_multiple(parameters, t):
    """[B]: alpha0 + beta sin(2 pi f t + phi) with parameters = [[alpha0, beta, f, phi], ...]."""
    t = _time(t)
    a0, beta, f, phi = (torch.stack([_as(x, t).reshape(()) for x in _col(parameters, i)]) for i in range(4))
    return a0 + beta * torch.sin(_TWO_PI * f * t + phi)


_CLOSED_FORM = {
    "displacement": _cf_displacement_multi,           # the single-body laws are the B = 1 case
    "rotation": _cf_rotation_multi,
    "displacement_multiple": _cf_displacement_multi,
    "rotation_multiple": _cf_rotation_multi,
    "Displacement_Foil_Fourier_Dotted_Mutliple": _cf_displacement_fourier,
    "rotation_Foil_Fourier_Dotted_Mutliple": _cf_rotation_fourier,
    "rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED": _cf_rotation_fourier_normalized,
}
