"""Immersed-boundary interpolation, forcing and spreading.  ORACLE (tests only).

Restates:
  * jax_ib/base/convolution_functions.py:5-7    delta_approx_logistjax (Gaussian, sigma = grid step)
  * jax_ib/base/convolution_functions.py:11-37  new_surf_fn      (E->L interpolation, dense sum)
  * jax_ib/base/IBM_Force.py:20-87              IBM_force_GENERAL (markers, rigid velocity,
                                                direct forcing, arc length, L->E spreading)
  * jax_ib/base/IBM_Force.py:89-115             IBM_Multiple_NEW, calc_IBM_force_NEW_MULTIPLE

`ndev = 1` semantics: the reference's pmap drops `M mod device_count` markers
(IBM_Force.py:77-78); with one device nothing is dropped.

Both a dense form (every marker against the whole grid, exactly what the reference
does) and a windowed form (2R x 2R cells around floor(xp/dx - offset), the form
the CUDA kernels compute) are provided; tests/test_oracle_identities.py measures the gap.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional, Sequence

import numpy as np

from .field import Var
from . import kinematics as kin


def gaussian_delta(x, x0, w):
    """convolution_functions.py:5-7: 1/(w sqrt(2 pi)) * exp(-0.5 ((x - x0)/w)^2)."""
    t = np.asarray(x).dtype.type if np.asarray(x).dtype.kind == "f" else np.asarray(x0).dtype.type
    w = t(w)
    pref = t(1) / (w * np.sqrt(t(2 * np.pi)))
    return pref * np.exp(t(-0.5) * ((x - x0) / w) ** 2)


def marker_cell_index(xp, yp, grid, offset):
    """The cell whose mesh point lies at or just below the marker: floor((xp - lower)/dx - offset).  The
    reference writes xp/dx - offset (IBM_Force.py:35, where the value is never used: its sums are dense) for
    domains that start at 0; subtracting a zero origin is exact, so both forms give the same index there."""
    t = xp.dtype.type
    si = (xp - t(grid.domain[0][0])) / t(grid.step[0]) - t(offset[0])
    sj = (yp - t(grid.domain[1][0])) / t(grid.step[1]) - t(offset[1])
    return np.floor(si).astype(np.int32), np.floor(sj).astype(np.int32)


def interpolate_dense(field: Var, xp, yp):
    """convolution_functions.py:11-37 with ndev = 1: U_k = sum_ij f * d(xp-X) * d(yp-Y) * dx * dy."""
    d = field.data
    t = d.dtype.type
    X, Y = field.grid.mesh(field.offset, d.dtype)
    dx, dy = field.grid.step
    out = np.empty(len(xp), dtype=d.dtype)
    for k in range(len(xp)):
        out[k] = np.sum(d * gaussian_delta(xp[k], X, dx) * gaussian_delta(yp[k], Y, dy) * t(dx) * t(dy))
    return out


def _window(i0, n, R):
    lo = max(i0 - R + 1, 0)
    hi = min(i0 + R, n - 1)
    return lo, hi + 1


def interpolate_windowed(field: Var, xp, yp, R=6):
    """Same sum restricted to the 2R x 2R cells i in [i_k-R+1, i_k+R] (clipped to the grid;
    the reference takes no periodic image: raw xp - X)."""
    d = field.data
    t = d.dtype.type
    ax, ay = field.grid.axes(field.offset, d.dtype)
    dx, dy = field.grid.step
    ik, jk = marker_cell_index(xp, yp, field.grid, field.offset)
    out = np.zeros(len(xp), dtype=d.dtype)
    for k in range(len(xp)):
        a, b = _window(int(ik[k]), d.shape[0], R)
        c, e = _window(int(jk[k]), d.shape[1], R)
        if a >= b or c >= e:
            continue
        gx = gaussian_delta(xp[k], ax[a:b], dx)[:, None]
        gy = gaussian_delta(yp[k], ay[c:e], dy)[None, :]
        out[k] = np.sum(d[a:b, c:e] * gx * gy * t(dx) * t(dy))
    return out


def spread_dense(F, xp, yp, dS, grid, offset, dtype=np.float32):
    """IBM_Force.py:66-87 with ndev = 1: f_ij = sum_k F_k * d(sqrt((xp-X)^2+(yp-Y)^2); dx) * dS_k."""
    X, Y = grid.mesh(offset, dtype)
    out = np.zeros(grid.shape, dtype=dtype)
    for k in range(len(xp)):
        r = np.sqrt((xp[k] - X) ** 2 + (yp[k] - Y) ** 2)
        out += F[k] * gaussian_delta(r, 0, grid.step[0]) * dS[k]
    return out


def spread_windowed(F, xp, yp, dS, grid, offset, R=6, dtype=np.float32):
    ax, ay = grid.axes(offset, dtype)
    ik, jk = marker_cell_index(xp, yp, grid, offset)
    out = np.zeros(grid.shape, dtype=dtype)
    for k in range(len(xp)):
        a, b = _window(int(ik[k]), grid.shape[0], R)
        c, e = _window(int(jk[k]), grid.shape[1], R)
        if a >= b or c >= e:
            continue
        r = np.sqrt((xp[k] - ax[a:b, None]) ** 2 + (yp[k] - ay[None, c:e]) ** 2)
        out[a:b, c:e] += F[k] * gaussian_delta(r, 0, grid.step[0]) * dS[k]
    return out


@dataclasses.dataclass
class Bodies:
    """particle (particle_class.py:106-139): centres [B,2] + per-body parameters; the shape and the
    two kinematic laws are static.  `shape_fn(geometry_param_b) -> (x0, y0)` body-frame markers;
    `displacement`/`rotation` are (value, derivative) pairs taking ([param_b...], t)."""

    centers: np.ndarray
    geometry: Sequence
    displacement_param: Sequence
    rotation_param: Sequence
    shape_fn: Callable
    displacement: tuple
    rotation: tuple

    def replace_centers(self, centers) -> "Bodies":
        return dataclasses.replace(self, centers=centers)


def body_markers(bodies: Bodies, b: int, t, dtype=np.float32):
    """IBM_Force.py:29-34: rotate body-frame markers by rotation(t), translate to the centre."""
    x0, y0 = bodies.shape_fn(bodies.geometry[b])
    x0 = np.asarray(x0, dtype=dtype)
    y0 = np.asarray(y0, dtype=dtype)
    theta = np.asarray(bodies.rotation[0]([bodies.rotation_param[b]], t, dtype), dtype=dtype).reshape(-1)[0]
    c, s = np.cos(theta), np.sin(theta)
    xc, yc = (dtype(v) for v in bodies.centers[b])
    xp = x0 * c - y0 * s + xc
    yp = x0 * s + y0 * c + yc
    return xp, yp


def ibm_force_one_body(field: Var, Xi: int, bodies: Bodies, b: int, dt, R: Optional[int] = None):
    """IBM_Force.py:20-87 for body `b`, component `Xi`; returns the raw force array."""
    d = field.data
    dtype = d.dtype.type
    t = field.bc.time_stamp
    xp, yp = body_markers(bodies, b, t, dtype)
    xc, yc = (dtype(v) for v in bodies.centers[b])
    U_interp = interpolate_dense(field, xp, yp) if R is None else interpolate_windowed(field, xp, yp, R)
    r = -(yp - yc) if Xi == 0 else (xp - xc)  # :39-42
    U0 = np.asarray(bodies.displacement[1]([bodies.displacement_param[b]], t, dtype), dtype=dtype).reshape(2, -1)[:, 0]
    omega = np.asarray(bodies.rotation[1]([bodies.rotation_param[b]], t, dtype), dtype=dtype).reshape(-1)[0]
    UP = U0[Xi] + omega * r  # :44-47
    force = (UP - U_interp) / dtype(dt)  # :50
    dxl = np.roll(xp, -1) - xp  # :59-63
    dyl = np.roll(yp, -1) - yp
    dS = np.sqrt(dxl ** 2 + dyl ** 2)
    if R is None:
        return spread_dense(force, xp, yp, dS, field.grid, field.offset, d.dtype)
    return spread_windowed(force, xp, yp, dS, field.grid, field.offset, R, d.dtype)


def ibm_force(velocity, bodies: Bodies, dt, R: Optional[int] = None):
    """IBM_Force.py:89-115: sum over bodies, per component (Xi = 0 on u, Xi = 1 on v)."""
    out = []
    for Xi, field in enumerate(velocity):
        f = np.zeros_like(field.data)
        for b in range(len(bodies.centers)):
            f = f + ibm_force_one_body(field, Xi, bodies, b, dt, R)
        out.append(f)
    return tuple(out)


# ---------------------------------------------------------------- the reference's marker sharding, on host threads
def ibm_force_threads(velocity, bodies: Bodies, dt, nthreads: int):
    """Dense ibm_force (R = None) with the MARKERS of every body sharded over `nthreads` host threads and the grid
    shared -- the reference's own parallelisation (jax.pmap over marker shards, grid replicated:
    IBM_Force.py:77-87, convolution_functions.py:28-35), with ndev = 1 semantics (no marker dropped).  NumPy
    releases the GIL inside its array kernels, so the shards run concurrently.  Used by bench.py's reference arm;
    partial sums are added in shard order, so the result differs from `ibm_force` only by summation order."""
    from concurrent.futures import ThreadPoolExecutor

    out = []
    with ThreadPoolExecutor(max_workers=max(1, nthreads)) as pool:
        for Xi, field in enumerate(velocity):
            d = field.data
            dtype = d.dtype.type
            t = field.bc.time_stamp
            f = np.zeros_like(d)
            for b in range(len(bodies.centers)):
                xp, yp = body_markers(bodies, b, t, dtype)
                xc, yc = (dtype(v) for v in bodies.centers[b])
                shards = [s for s in np.array_split(np.arange(len(xp)), max(1, nthreads)) if len(s)]
                U = np.concatenate(list(pool.map(lambda s: interpolate_dense(field, xp[s], yp[s]), shards)))
                r = -(yp - yc) if Xi == 0 else (xp - xc)
                U0 = np.asarray(bodies.displacement[1]([bodies.displacement_param[b]], t, dtype), dtype=dtype).reshape(2, -1)[:, 0]
                omega = np.asarray(bodies.rotation[1]([bodies.rotation_param[b]], t, dtype), dtype=dtype).reshape(-1)[0]
                force = (U0[Xi] + omega * r - U) / dtype(dt)
                dS = np.sqrt((np.roll(xp, -1) - xp) ** 2 + (np.roll(yp, -1) - yp) ** 2)
                parts = pool.map(lambda s: spread_dense(force[s], xp[s], yp[s], dS[s], field.grid, field.offset, d.dtype), shards)
                for part in parts:
                    f = f + part
            out.append(f)
    return tuple(out)


# ---------------------------------------------------------------- demo shapes (notebook cells)
def ellipse_shape(geometry_param, ntheta=150, dtype=np.float32):
    """Flapping_Demo.ipynb cell 5 foil_XY_ELLIPSE: 2*ntheta - 2 markers."""
    A, B = (dtype(v) for v in geometry_param[:2])
    xt = np.linspace(-A, A, ntheta, dtype=dtype)
    yt = B / A * np.sqrt(A ** 2 - xt ** 2)
    xt2 = np.linspace(A, -A, ntheta, dtype=dtype)[1:-1]
    yt2 = -B / A * np.sqrt(A ** 2 - xt2 ** 2)
    return np.append(xt, xt2), np.append(yt, yt2)


def circle_shape(geometry_param, ntheta=80, dtype=np.float32):
    """journal_bearing_demo.ipynb cell 6 foil_XY_circle."""
    A = dtype(geometry_param[0])
    theta = np.linspace(0, dtype(1.98 * np.pi), ntheta, dtype=dtype)
    return A * np.cos(theta), A * np.sin(theta)
