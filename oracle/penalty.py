"""Brinkman-penalisation mask and forcing.  ORACLE (tests only).

Restates:
  * jax_ib/penalty/util_funs.py:6-16     arbitrary_obstacle  (forcing = grad_p - kappa * u)
  * jax_ib/penalty/util_funs.py:26-73    project_particle    (polar angle, piecewise-linear R(theta))
  * jax_ib/penalty/util_funs.py:76-93    calc_perm           (kappa = smoothening_fn(R^2 - r^2, Know))
  * jax_ib/penalty/util_funs.py:95-123   perm_vmap_multiple_particles (sum over bodies)
  * jax_ib/penalty/parameteric_fns.py:3-26  param_ellipse / rose / snail / circle / rot_ellipse
"""
from __future__ import annotations

import numpy as np


def param_ellipse(g, theta):
    A, B = g[0], g[1]
    return A * B / np.sqrt((B * np.cos(theta)) ** 2 + (A * np.sin(theta)) ** 2)


def param_rose(g, theta):
    return g[0] * np.sin(g[1] * theta)


def param_snail(g, theta):
    return g[0] * theta


def param_circle(g, theta):
    return g[0] * np.ones_like(theta)


def polar_radius(grid, center, Rtheta, dtype=np.float32):
    """util_funs.py:26-61: per-cell polar angle about `center` (three-branch arctan), segment
    index theta // dtheta, linear interpolation of the tabulated R(theta).  Returns (Rfinal, theta)."""
    t = np.dtype(dtype).type
    xc, yc = (t(c) for c in center)
    X, Y = grid.mesh(grid.cell_center, dtype)
    Rtheta = np.asarray(Rtheta, dtype=dtype)
    n = Rtheta.size
    theta_tab = np.linspace(0, 2 * np.pi, n).astype(dtype)
    theta_next = np.roll(theta_tab, -1)
    R_next = np.roll(Rtheta, -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        at = np.arctan((Y - yc) / (X - xc))
    H = np.heaviside
    one, zero = t(1), t(0)
    th = at * H(X - xc, one) * H(Y - yc, zero)
    th = th + (at + t(np.pi)) * H(xc - X, one)
    th = th + (at + t(2 * np.pi)) * H(X - xc, zero) * H(yc - Y, zero)
    dtheta = t(2 * np.pi / (n - 1))
    idx = (th // dtheta).astype(np.int64)
    R0, R1, th0 = Rtheta[idx], R_next[idx], theta_tab[idx]
    slope = (R1 - R0) / (theta_next[idx] - th0)
    return R0 + slope * (th - th0), th


def calc_perm(grid, center, Rtheta, smoothening_fn, Know, dtype=np.float32):
    """util_funs.py:76-93: G = Rfinal^2 - r^2 ; kappa = smoothening_fn(G, Know)."""
    t = np.dtype(dtype).type
    X, Y = grid.mesh(grid.cell_center, dtype)
    Rfinal, _ = polar_radius(grid, center, Rtheta, dtype)
    G = Rfinal ** 2 - ((Y - t(center[1])) ** 2 + (X - t(center[0])) ** 2)
    return smoothening_fn(G, Know)


def tanh_smoothing(G, Know):
    """The smoothening the reference's comments suggest (util_funs.py:91): K/2 (1 + tanh(G / width))."""
    K, width = Know
    t = G.dtype.type
    return t(K) / t(2) * (t(1) + np.tanh(G / t(width)))


def perm_multiple(grid, centers, Rthetas, smoothening_fn, Know, dtype=np.float32):
    """util_funs.py:95-123."""
    out = np.zeros(grid.shape, dtype=dtype)
    for c, R in zip(centers, Rthetas):
        out = out + calc_perm(grid, c, R, smoothening_fn, Know, dtype)
    return out


def obstacle_forcing(pressure_gradient, permeability):
    """util_funs.py:6-16: forcing(v) = (px - kappa * u, py - kappa * v)."""

    def forcing(v):
        return tuple(v_i.data.dtype.type(g) * np.ones_like(v_i.data) - permeability * v_i.data
                     for g, v_i in zip(pressure_gradient, v))

    return forcing
