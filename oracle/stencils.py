"""Finite differences, face interpolation, advection and diffusion.  ORACLE (tests only).

Restates:
  * jax_ib/base/finite_differences.py:46-56,97-146   stencil_sum, backward/forward difference,
                                                     laplacian, divergence
  * jax_ib/base/interpolation.py:37-156              _linear_along_axis, linear, upwind
  * jax_ib/base/interpolation.py:159-304             lax_wendroff, van_leer_limiter, apply_tvd_limiter
  * jax_ib/base/advection.py:33-124,206-280,325-333  _advect_aligned, advect_general, advect_upwind,
                                                     advect_van_leer, advect_van_leer_using_limiters
  * jax_ib/base/diffusion.py:35-37                   diffuse
  * jax_ib/base/grids.py:511-517                     control_volume_offsets
  * jax_ib/base/equations.py:42-83                   navier_stokes_explicit_terms
"""
from __future__ import annotations

import numpy as np

from .field import Var, periodic_bc


def _s(x, like):
    """Round a Python scalar the way JAX weak typing does when it meets `like`."""
    return like.dtype.type(x)


# ----------------------------------------------------------------------------- differences
def forward_difference(u: Var, axis: int) -> np.ndarray:
    """finite_differences.py:119-126: (u[+1] - u) / h, true division."""
    return (u.shift(+1, axis).data + (-u.data)) / _s(u.grid.step[axis], u.data)


def backward_difference(u: Var, axis: int) -> np.ndarray:
    """finite_differences.py:97-104: (u - u[-1]) / h."""
    return (u.data + (-u.shift(-1, axis).data)) / _s(u.grid.step[axis], u.data)


def laplacian(u: Var) -> np.ndarray:
    """finite_differences.py:129-136."""
    scales = np.square(1 / np.array(u.grid.step, dtype=u.data.dtype))
    result = _s(-2, u.data) * u.data * np.sum(scales)
    for axis in range(2):
        result = result + (u.shift(-1, axis).data + u.shift(+1, axis).data) * scales[axis]
    return result


def divergence(v) -> np.ndarray:
    """finite_differences.py:139-146: sum of backward differences (lands on cell centres)."""
    return backward_difference(v[0], 0) + backward_difference(v[1], 1)


def diffuse(c: Var, nu: float) -> np.ndarray:
    """diffusion.py:35-37."""
    return _s(nu, c.data) * laplacian(c)


# ----------------------------------------------------------------------------- interpolation
def control_volume_offsets(c: Var):
    """grids.py:511-517."""
    return tuple(tuple(o + 0.5 if i == j else o for i, o in enumerate(c.offset)) for j in range(2))


def _linear_along_axis(c: Var, offset: float, axis: int) -> Var:
    """interpolation.py:37-65."""
    delta = offset - c.offset[axis]
    if delta == 0:
        return c
    new_offset = tuple(offset if j == axis else o for j, o in enumerate(c.offset))
    if int(delta) == delta:
        return c.with_data(c.shift(int(delta), axis).data, new_offset)
    lo, hi = int(np.floor(delta)), int(np.ceil(delta))
    w_lo = hi - delta
    w_hi = 1.0 - w_lo
    data = _s(w_lo, c.data) * c.shift(lo, axis).data + _s(w_hi, c.data) * c.shift(hi, axis).data
    return c.with_data(data, new_offset)


def linear(c: Var, offset, v=None, dt=None) -> Var:
    """interpolation.py:68-94."""
    out = c
    for a, o in enumerate(offset):
        out = _linear_along_axis(out, o, a)
    return out


def _interp_axis(c: Var, offset):
    axes = [a for a, (cur, tgt) in enumerate(zip(c.offset, offset)) if cur != tgt]
    if len(axes) != 1:
        raise ValueError("offsets must differ in exactly one entry")
    return axes[0]


def upwind(c: Var, offset, v, dt=None) -> Var:
    """interpolation.py:97-156: where(u > 0, c[floor], c[ceil]); result carries periodic BC."""
    if tuple(c.offset) == tuple(offset):
        return c
    axis = _interp_axis(c, offset)
    u = v[axis]
    delta = u.offset[axis] - c.offset[axis]
    if int(delta) == delta:
        return c.with_data(c.shift(int(delta), axis).data, tuple(offset))
    lo, hi = int(np.floor(delta)), int(np.ceil(delta))
    data = np.where(u.data > 0, c.shift(lo, axis).data, c.shift(hi, axis).data)
    return Var(data, tuple(offset), c.grid, periodic_bc())


def lax_wendroff(c: Var, offset, v, dt) -> Var:
    """interpolation.py:159-222."""
    if tuple(c.offset) == tuple(offset):
        return c
    axis = _interp_axis(c, offset)
    u = v[axis]
    delta = u.offset[axis] - c.offset[axis]
    lo, hi = int(np.floor(delta)), int(np.ceil(delta))
    courant = _s(dt / c.grid.step[axis], c.data) * u.data
    c_lo, c_hi = c.shift(lo, axis).data, c.shift(hi, axis).data
    half, one = _s(0.5, c.data), _s(1, c.data)
    pos = c_lo + half * (one - courant) * (c_hi - c_lo)
    neg = c_hi - half * (one + courant) * (c_hi - c_lo)
    return Var(np.where(u.data > 0, pos, neg), tuple(offset), c.grid, periodic_bc())


def _safe_div(x, y):
    """interpolation.py:225-227."""
    return x / np.where(y != 0, y, y.dtype.type(1))


def van_leer_limiter(r):
    """interpolation.py:230-232."""
    return np.where(r > 0, _safe_div(r.dtype.type(2) * r, r.dtype.type(1) + r), r.dtype.type(0))


def tvd_van_leer(c: Var, offset, v, dt) -> Var:
    """interpolation.py:235-304 apply_tvd_limiter(lax_wendroff, van_leer_limiter)."""
    for axis, axis_offset in enumerate(offset):
        target = tuple(o if i != axis else axis_offset for i, o in enumerate(c.offset))
        if target != tuple(c.offset):
            if target[axis] - c.offset[axis] != 0.5:
                raise NotImplementedError("tvd_interpolation only supports forward interpolation")
            c_low = upwind(c, offset, v, dt)
            c_high = lax_wendroff(c, offset, v, dt)
            c_left = c.shift(-1, axis).data
            c_right = c.shift(1, axis).data
            c_next = c.shift(2, axis).data
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                r_pos = _safe_div(c.data - c_left, c_right - c.data)
                r_neg = _safe_div(c_next - c_right, c_right - c.data)
                phi = np.where(v[axis].data > 0, van_leer_limiter(r_pos), van_leer_limiter(r_neg))
            data = c_low.data - (c_low.data - c_high.data) * phi
            c = c.with_data(data, target)
    return c


# ----------------------------------------------------------------------------- advection
def _advect_aligned(cs, v) -> np.ndarray:
    """advection.py:33-75: -div(c*u).  Each flux component inherits the BC of the INTERPOLATED variable it was
    built from (advection.py:72-74), not the BC of the advected scalar: `upwind` hands back a periodic variable
    (interpolation.py:152-156), so with plain upwind advection the flux wraps around even across a wall, while the
    TVD and van Leer paths keep the scalar's own BC."""
    flux = tuple(Var(c.data * u.data, c.offset, c.grid, c.bc) for c, u in zip(cs, v))
    return -divergence(flux)


def advect_general(c: Var, v, u_interp, c_interp, dt=None) -> np.ndarray:
    """advection.py:78-110."""
    targets = control_volume_offsets(c)
    aligned_v = tuple(u_interp(u, t, v, dt) for u, t in zip(v, targets))
    aligned_c = tuple(c_interp(c, t, aligned_v, dt) for t in targets)
    return _advect_aligned(aligned_c, aligned_v)


def advect_upwind(c: Var, v, dt=None) -> np.ndarray:
    """advection.py:120-124."""
    return advect_general(c, v, linear, upwind, dt)


def advect_van_leer_using_limiters(c: Var, v, dt) -> np.ndarray:
    """advection.py:325-333 (the default `convect`, equations.py:55-58)."""
    return advect_general(c, v, linear, tvd_van_leer, dt)


def advect_van_leer(c: Var, v, dt) -> np.ndarray:
    """advection.py:206-280."""
    targets = control_volume_offsets(c)
    aligned_v = tuple(linear(u, t) for u, t in zip(v, targets))
    flux = []
    d = c.data
    half, one, two, zero = _s(0.5, d), _s(1, d), _s(2, d), _s(0, d)
    for axis, (u, h) in enumerate(zip(aligned_v, c.grid.step)):
        c_center = d
        c_left = c.shift(-1, axis).data
        c_right = c.shift(+1, axis).data
        upwind_flux = np.where(u.data > 0, u.data * c_center, u.data * c_right)
        diffs_prod = two * (c_right - c_center) * (c_center - c_left)
        neighbor_diff = c_right - c_left
        safe = diffs_prod > 0
        forward_correction = np.where(safe, diffs_prod / np.where(safe, neighbor_diff, one), zero)
        # the correction is shifted with the *face velocity's* BC object (:267-270)
        backward_correction = Var(forward_correction, u.offset, u.grid, u.bc).shift(+1, axis).data
        abs_velocity = np.abs(u.data)
        courant = _s(dt / h, d) * abs_velocity
        pre_factor = half * (one - courant) * abs_velocity
        flux_correction = pre_factor * np.where(u.data > 0, forward_correction, backward_correction)
        flux.append(Var(upwind_flux + flux_correction, targets[axis], c.grid, c.bc))
    return -divergence(tuple(flux))


CONVECT = {
    "upwind": advect_upwind,
    "van_leer": advect_van_leer,
    "van_leer_limiters": advect_van_leer_using_limiters,
}


def explicit_terms(v, nu, dt, convect="upwind", forcing=None, density=1.0):
    """equations.py:42-83: dv/dt = convect(v) + (nu/rho) * laplacian(v) [+ forcing / rho].
    `forcing(v)` returns a tuple of arrays."""
    adv = tuple(CONVECT[convect](u, v, dt) for u in v)
    out = tuple(a + diffuse(u, nu / density) for a, u in zip(adv, v)) if nu is not None else adv
    if forcing is not None:
        out = tuple(o + f / _s(density, o) for o, f in zip(out, forcing(v)))
    return out
