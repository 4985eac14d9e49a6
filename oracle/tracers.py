"""Passive-tracer coupling (jax-md Brownian integrator driven by the CFD velocity).  ORACLE (tests only).

Restates:
  * jax_ib/MD/CFD_MD_coupling.py:4-27        surface_fn_jax, interpolate_pbc, custom_force_fn_pbc
  * jax_ib/MD/simulate.py:41-99              brownian_cfd (init_fn / apply_fn)
  * jax_ib/MD/interaction_potential.py:6-22  harmonic_morse, harmonic_morse_pair
and, un-vendored (PyPI jax-md, unpinned):
  * jax.scipy.ndimage.map_coordinates(order=1, mode='constant', cval=0)
  * jax_md.smap.pair (sum over unordered pairs, per-species parameter matrices),
    jax_md.quantity.force (F = -dE/dR), jax_md.space.periodic(side, wrapped=False)
    (displacement = minimum image, shift does not wrap).

Only kT = 0 dynamics are deterministic without JAX's threefry PRNG; for kT > 0 the caller
supplies the normal deviates `xi` explicitly.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np


def map_coordinates_linear(a: np.ndarray, ci: np.ndarray, cj: np.ndarray) -> np.ndarray:
    """Bilinear interpolation at fractional indices; taps outside the array contribute 0."""
    t = a.dtype.type
    i0 = np.floor(ci)
    j0 = np.floor(cj)
    fi = (ci - i0).astype(a.dtype)
    fj = (cj - j0).astype(a.dtype)
    i0 = i0.astype(np.int64)
    j0 = j0.astype(np.int64)
    out = np.zeros(ci.shape, dtype=a.dtype)
    for di, wi in ((0, t(1) - fi), (1, fi)):
        for dj, wj in ((0, t(1) - fj), (1, fj)):
            ii, jj = i0 + di, j0 + dj
            ok = (ii >= 0) & (ii < a.shape[0]) & (jj >= 0) & (jj < a.shape[1])
            val = np.where(ok, a[np.clip(ii, 0, a.shape[0] - 1), np.clip(jj, 0, a.shape[1] - 1)], t(0))
            out = out + wi * wj * val
    return out


def interpolate_field(field, R: np.ndarray) -> np.ndarray:
    """CFD_MD_coupling.py:7-19: index = x/dx - offset, bilinear, zero outside (no wrap)."""
    t = field.data.dtype.type
    ci = R[:, 0] / t(field.grid.step[0]) - t(field.offset[0])
    cj = R[:, 1] / t(field.grid.step[1]) - t(field.offset[1])
    return map_coordinates_linear(field.data, ci, cj)


def cfd_force(velocity, R: np.ndarray) -> np.ndarray:
    """CFD_MD_coupling.py:22-27: [u(R), v(R)] * 1.0."""
    return np.stack([interpolate_field(velocity[0], R), interpolate_field(velocity[1], R)], axis=-1)


def harmonic_morse(dr, h, D0, alpha, r0, k):
    """interaction_potential.py:6-11."""
    return np.where(dr < r0, h * k * (dr - r0) ** 2 - D0,
                    D0 * (np.exp(-2.0 * alpha * (dr - r0)) - 2.0 * np.exp(-alpha * (dr - r0))))


def harmonic_morse_force(R, species, h, r0, box, D0=0.0, alpha=10.0, k=1.0):
    """-d/dR of sum over unordered pairs of harmonic_morse(|minimum-image displacement|) with
    per-species-pair h and r0 (interaction_potential.py:15-22 through smap.pair + quantity.force).
    Analytic gradient: dU/dr = 2 h k (r - r0) for r < r0, else D0*(-2a e^{-2a(r-r0)} + 2a e^{-a(r-r0)})."""
    t = R.dtype.type
    d = R[:, None, :] - R[None, :, :]
    box = np.asarray(box, dtype=R.dtype)
    d = d - box * np.round(d / box)  # space.periodic displacement (minimum image)
    r = np.sqrt((d ** 2).sum(-1))
    hh = np.asarray(h, dtype=R.dtype)[species[:, None], species[None, :]]
    rr = np.asarray(r0, dtype=R.dtype)[species[:, None], species[None, :]]
    e1 = np.exp(-t(2) * t(alpha) * (r - rr))
    e2 = np.exp(-t(alpha) * (r - rr))
    dU = np.where(r < rr, t(2) * hh * t(k) * (r - rr), t(D0) * (-t(2) * t(alpha) * e1 + t(2) * t(alpha) * e2))
    np.fill_diagonal(dU, 0)
    safe_r = np.where(r > 0, r, t(1))
    return (-(dU / safe_r)[:, :, None] * d * (r > 0)[:, :, None]).sum(axis=1)


@dataclasses.dataclass
class BrownianState:
    position: np.ndarray
    mass: float = 1.0


def brownian_step(state: BrownianState, velocity, dt, gamma=1.0, kT=0.0, pair_force=None,
                  is_mobile: Optional[np.ndarray] = None, xi: Optional[np.ndarray] = None) -> BrownianState:
    """MD/simulate.py:81-97: dR = F dt nu + sqrt(2 kT dt nu) xi, nu = 1/(mass gamma);
    R = where(is_mobile, R + dR, R) (the notebooks' masked_shift; shift does not wrap)."""
    R = state.position
    t = R.dtype.type
    F = cfd_force(velocity, R)
    if pair_force is not None:
        F = pair_force(R) + F
    nu = t(1) / (t(state.mass) * t(gamma))
    dR = F * t(dt) * nu
    if kT != 0.0:
        if xi is None:
            raise ValueError("kT > 0 needs explicit normal deviates xi (JAX threefry is not restated)")
        dR = dR + np.sqrt(t(2) * t(kT) * t(dt) * nu) * xi.astype(R.dtype)
    newR = R + dR
    if is_mobile is not None:
        newR = np.where(np.asarray(is_mobile, dtype=bool)[:, None], newR, R)
    return BrownianState(newR, state.mass)
