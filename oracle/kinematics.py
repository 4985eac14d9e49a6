"""Prescribed rigid-body kinematics and their time derivatives.  ORACLE (tests only).

Restates jax_ib/base/kinematics.py:5-105.  The reference differentiates these with
`jax.jacrev` (IBM_Force.py:101-104, particle_motion.py:47-52); JAX is not available
here, so each law is paired with its closed-form derivative, evaluated in the same
precision.  Every function takes the reference's parameter layout and a scalar `t`
and returns NumPy values in `dtype` (float32 by default).
"""
from __future__ import annotations

import numpy as np

_F = np.float32


def _t(x, dtype):
    return np.dtype(dtype).type(x)


# --------------------------------------------------------------- Flapping_Demo (single harmonic)
def displacement(parameters, t, dtype=_F):
    """kinematics.py:5-7: [A0/2 cos(2 pi f t), 0]; parameters = [[A0, f]]."""
    A0, f = (_t(x, dtype) for x in parameters[0])
    t = _t(t, dtype)
    two_pi = _t(2 * np.pi, dtype)
    return np.array([A0 / _t(2, dtype) * np.cos(two_pi * f * t), _t(0, dtype)], dtype=dtype)


def displacement_dot(parameters, t, dtype=_F):
    """d/dt of `displacement` (what jacrev returns)."""
    A0, f = (_t(x, dtype) for x in parameters[0])
    t = _t(t, dtype)
    two_pi = _t(2 * np.pi, dtype)
    return np.array([-(A0 / _t(2, dtype)) * np.sin(two_pi * f * t) * (two_pi * f), _t(0, dtype)], dtype=dtype)


def rotation(parameters, t, dtype=_F):
    """kinematics.py:10-12: alpha0 + beta sin(2 pi f t + phi); parameters = [[alpha0, beta, f, phi]]."""
    a0, beta, f, phi = (_t(x, dtype) for x in parameters[0])
    t = _t(t, dtype)
    return a0 + beta * np.sin(_t(2 * np.pi, dtype) * f * t + phi)


def rotation_dot(parameters, t, dtype=_F):
    a0, beta, f, phi = (_t(x, dtype) for x in parameters[0])
    t = _t(t, dtype)
    two_pi = _t(2 * np.pi, dtype)
    return beta * np.cos(two_pi * f * t + phi) * (two_pi * f)


# --------------------------------------------------------------- Fourier-series multi-body laws
def _unpack(parameters, n, dtype):
    cols = list(zip(*parameters))
    return [np.asarray(cols[i], dtype=dtype) for i in range(n)]


def _series(parameters, t, dtype, n_fields=6, p_index=5):
    fields = _unpack(parameters, n_fields, dtype)
    alpha0, f, phi, alpha, beta = fields[:5]
    p = fields[p_index]
    alpha = alpha.reshape(len(alpha0), -1)
    beta = beta.reshape(len(alpha0), -1)
    nb, nt = alpha.shape
    freq = np.tile(np.arange(1, nt + 1, dtype=dtype), (nb, 1))
    return alpha0, f.reshape(nb, 1), phi.reshape(nb, 1), alpha, beta, p, freq


def displacement_fourier(parameters, t, dtype=_F):
    """kinematics.py:14-38 Displacement_Foil_Fourier_Dotted_Mutliple: returns [2, B]."""
    alpha0, f, phi, alpha, beta, p, freq = _series(parameters, t, dtype)
    t = _t(t, dtype)
    inside = _t(2 * np.pi, dtype) * t * freq * f + phi
    a1 = (alpha * np.sin(inside)).sum(axis=1) + p * (beta * np.cos(inside)).sum(axis=1)
    return np.array([-alpha0 * t, a1], dtype=dtype)


def displacement_fourier_dot(parameters, t, dtype=_F):
    alpha0, f, phi, alpha, beta, p, freq = _series(parameters, t, dtype)
    t = _t(t, dtype)
    w = _t(2 * np.pi, dtype) * freq * f
    inside = _t(2 * np.pi, dtype) * t * freq * f + phi
    a1 = (alpha * np.cos(inside) * w).sum(axis=1) - p * (beta * np.sin(inside) * w).sum(axis=1)
    return np.array([-alpha0 * np.ones_like(a1), a1], dtype=dtype)


def rotation_fourier(parameters, t, dtype=_F):
    """kinematics.py:41-65 rotation_Foil_Fourier_Dotted_Mutliple: returns [B]."""
    alpha0, f, phi, alpha, beta, p, freq = _series(parameters, t, dtype)
    t = _t(t, dtype)
    inside = _t(2 * np.pi, dtype) * t * freq * f + phi
    a1 = (alpha * np.sin(inside)).sum(axis=1) + p * (beta * np.cos(inside)).sum(axis=1)
    return alpha0 * t + a1


def rotation_fourier_dot(parameters, t, dtype=_F):
    alpha0, f, phi, alpha, beta, p, freq = _series(parameters, t, dtype)
    t = _t(t, dtype)
    w = _t(2 * np.pi, dtype) * freq * f
    inside = _t(2 * np.pi, dtype) * t * freq * f + phi
    return alpha0 + (alpha * np.cos(inside) * w).sum(axis=1) - p * (beta * np.sin(inside) * w).sum(axis=1)


def _norm_parts(parameters, t, dtype):
    alpha0, f, phi, alpha, beta, p, freq = _series(parameters, t, dtype, n_fields=7, p_index=6)
    theta_av = _unpack(parameters, 7, dtype)[5]
    two_pi = _t(2 * np.pi, dtype)
    in2 = two_pi * freq + phi
    a2 = (alpha * np.sin(in2)).sum(axis=1) + p * (beta * np.cos(in2)).sum(axis=1)
    in3 = two_pi * freq * _t(0, dtype) + phi
    a3 = (alpha * np.sin(in3)).sum(axis=1) + p * (beta * np.cos(in3)).sum(axis=1)
    return alpha0, f, phi, alpha, beta, p, freq, theta_av, a2, a3


def rotation_fourier_normalized(parameters, t, dtype=_F):
    """kinematics.py:67-105 rotation_Foil_Fourier_Dotted_Mutliple_NORMALIZED: returns [B]."""
    alpha0, f, phi, alpha, beta, p, freq, theta_av, a2, a3 = _norm_parts(parameters, t, dtype)
    t = _t(t, dtype)
    inside = _t(2 * np.pi, dtype) * t * freq * f + phi
    a1 = (alpha * np.sin(inside)).sum(axis=1) + p * (beta * np.cos(inside)).sum(axis=1)
    return theta_av * (alpha0 * t + a1) / (alpha0 + a2 - a3)


def rotation_fourier_normalized_dot(parameters, t, dtype=_F):
    alpha0, f, phi, alpha, beta, p, freq, theta_av, a2, a3 = _norm_parts(parameters, t, dtype)
    t = _t(t, dtype)
    w = _t(2 * np.pi, dtype) * freq * f
    inside = _t(2 * np.pi, dtype) * t * freq * f + phi
    d1 = (alpha * np.cos(inside) * w).sum(axis=1) - p * (beta * np.sin(inside) * w).sum(axis=1)
    return theta_av * (alpha0 + d1) / (alpha0 + a2 - a3)


# name -> (value, derivative).  Per-body evaluation (`[param_b]`) as IBM_Force.py:101-104 does.
DISPLACEMENT = {
    "harmonic": (displacement, displacement_dot),
    "fourier": (displacement_fourier, displacement_fourier_dot),
}
ROTATION = {
    "harmonic": (rotation, rotation_dot),
    "fourier": (rotation_fourier, rotation_fourier_dot),
    "fourier_normalized": (rotation_fourier_normalized, rotation_fourier_normalized_dot),
}


# ---- multi-body harmonic laws (vectorised form of `displacement` / `rotation`; the reference's own
# versions take a single body).  Used by the multi-body synthetic workloads.
def displacement_multi(parameters, t, dtype=_F):
    cols = _unpack(parameters, 2, dtype)
    A0, f = cols
    t = _t(t, dtype)
    x = A0 / _t(2, dtype) * np.cos(_t(2 * np.pi, dtype) * f * t)
    return np.array([x, np.zeros_like(x)], dtype=dtype)


def displacement_multi_dot(parameters, t, dtype=_F):
    A0, f = _unpack(parameters, 2, dtype)
    t = _t(t, dtype)
    two_pi = _t(2 * np.pi, dtype)
    x = -(A0 / _t(2, dtype)) * np.sin(two_pi * f * t) * (two_pi * f)
    return np.array([x, np.zeros_like(x)], dtype=dtype)


def rotation_multi(parameters, t, dtype=_F):
    a0, beta, f, phi = _unpack(parameters, 4, dtype)
    t = _t(t, dtype)
    return a0 + beta * np.sin(_t(2 * np.pi, dtype) * f * t + phi)


def rotation_multi_dot(parameters, t, dtype=_F):
    a0, beta, f, phi = _unpack(parameters, 4, dtype)
    t = _t(t, dtype)
    two_pi = _t(2 * np.pi, dtype)
    return beta * np.cos(two_pi * f * t + phi) * (two_pi * f)


DISPLACEMENT["harmonic_multi"] = (displacement_multi, displacement_multi_dot)
ROTATION["harmonic_multi"] = (rotation_multi, rotation_multi_dot)
