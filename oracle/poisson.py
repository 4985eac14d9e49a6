"""Pressure projection by fast diagonalisation.  ORACLE (tests only).

Restates:
  * jax_ib/base/array_utils.py:168-182,254-298  laplacian_matrix, laplacian_matrix_neumann,
                                                laplacian_matrix_w_boundaries
  * jax_ib/base/pressure.py:33-55               _rhs_transform
  * jax_ib/base/pressure.py:58-91               projection_and_update_pressure
  * jax_ib/base/pressure.py:94-154              solve_fast_diag, solve_fast_diag_moving_wall,
                                                solve_fast_diag_Far_Field
and, because it is NOT under /root/reference (PyPI `jax-cfd`, unpinned, reference
setup.py:5; called at pressure.py:105-107,127-129,151-153), the published algorithm of
  * jax_cfd.base.fast_diagonalization.transform / pseudoinverse:
      cutoff = 10 * finfo(dtype).eps ; func(l) = where(|l| > cutoff, 1/l, 0)
      'rfft'  : eigenvalues = fft(op[:, 0]) (leading axes) / rfft(op[:, 0]) (last axis),
                outer-summed in float64, func applied, cast to the complex dtype;
                apply = irfftn(diag * rfftn(rhs)).astype(dtype)
      'matmul': w, V = eigh(op) per axis (float64 on host, V cast to dtype);
                out = rhs; for V: out = tensordot(out, V, axes=(0, 0));
                out *= func(outer-sum w); for V: out = tensordot(out, V, axes=(0, 1))
      implementation=None -> 'rfft' if circulant (and the leading sizes are even) else 'matmul'.
"""
from __future__ import annotations

import functools

import numpy as np
import scipy.fft
import scipy.linalg

from .field import Var, DIRICHLET, NEUMANN, PERIODIC, all_periodic, pressure_bc_from_velocity
from . import stencils


# ------------------------------------------------------------------ 1-D operators
def laplacian_matrix(size: int, step: float) -> np.ndarray:
    """array_utils.py:168-173 (periodic, circulant)."""
    col = np.zeros(size)
    col[0] = -2 / step ** 2
    col[1] = col[-1] = 1 / step ** 2
    return scipy.linalg.circulant(col)


def laplacian_matrix_neumann(size: int, step: float) -> np.ndarray:
    """array_utils.py:175-182 (homogeneous Neumann, cell centred)."""
    col = np.zeros(size)
    col[0] = -2 / step ** 2
    col[1] = 1 / step ** 2
    m = scipy.linalg.toeplitz(col)
    m[0, 0] = m[-1, -1] = -1 / step ** 2
    return m


def laplacian_matrix_w_boundaries(grid, offset, bc):
    """array_utils.py:254-298 (cell-centred offsets only, which is what pressure uses)."""
    mats = [laplacian_matrix(n, h) for n, h in zip(grid.shape, grid.step)]
    for axis in range(2):
        if np.isclose(offset[axis], 0.5):
            for i in (0, 1):
                kind = bc.types[axis][i]
                if kind in (NEUMANN, DIRICHLET):
                    delta = 1 / grid.step[axis] ** 2
                    idx = (0, 0) if i == 0 else (-1, -1)
                    mats[axis][idx] += delta if kind == NEUMANN else -delta
                    mats[axis][0, -1] = 0.0
                    mats[axis][-1, 0] = 0.0
    return mats


# ------------------------------------------------------------------ jax_cfd fast_diagonalization
def _inv_func(dtype):
    cutoff = 10 * np.finfo(dtype).eps
    return lambda lam: np.where(np.abs(lam) > cutoff, 1 / np.where(np.abs(lam) > cutoff, lam, 1), 0)


def pseudoinverse(operators, dtype, circulant: bool, implementation=None):
    """Returns apply(rhs) for the pseudo-inverse of sum_axis kron(..., op_axis, ...)."""
    dtype = np.dtype(dtype)
    func = _inv_func(dtype)
    if implementation is None:
        implementation = "rfft" if circulant else "matmul"
    if implementation == "rfft" and any(op.shape[0] % 2 for op in operators[:-1]):
        implementation = "matmul"
    if implementation == "rfft":
        eig = [np.fft.fft(op[:, 0]) for op in operators[:-1]] + [np.fft.rfft(operators[-1][:, 0])]
        diag = func(functools.reduce(np.add.outer, eig))
        cdtype = np.complex64 if dtype == np.float32 else np.complex128
        diag = diag.astype(cdtype)

        def apply(rhs):
            # scipy.fft keeps single precision for float32 input (pocketfft templated on float), as
            # jnp.fft / XLA do; numpy.fft would silently transform complex64 in double (measured:
            # rel-L2 2.5e-8 vs 1.1e-7 at n = 1024), which is NOT what the reference computes.
            axes = tuple(range(rhs.ndim))
            spec = scipy.fft.rfftn(rhs.astype(dtype), axes=axes)
            return scipy.fft.irfftn((diag * spec).astype(cdtype), s=rhs.shape, axes=axes).astype(dtype)

        return apply
    if implementation == "matmul":
        ws, vs = zip(*(np.linalg.eigh(op) for op in operators))
        diag = func(functools.reduce(np.add.outer, ws)).astype(dtype)
        vs = [v.astype(dtype) for v in vs]

        def apply(rhs):
            out = rhs.astype(dtype)
            for v in vs:
                out = np.tensordot(out, v, axes=(0, 0))
            out = out * diag
            for v in vs:
                out = np.tensordot(out, v, axes=(0, 1))
            return out.astype(dtype)

        return apply
    raise NotImplementedError(implementation)


# ------------------------------------------------------------------ solves (pressure.py)
def solve_fast_diag(v, implementation=None):
    """pressure.py:94-108: periodic x periodic."""
    if not all_periodic(*v):
        raise ValueError("solve_fast_diag() expects periodic velocity BC")
    grid = v[0].grid
    rhs = stencils.divergence(v)
    ops = [laplacian_matrix(n, h) for n, h in zip(grid.shape, grid.step)]
    return pseudoinverse(ops, rhs.dtype, circulant=True, implementation=implementation)(rhs)


def solve_fast_diag_moving_wall(v, implementation="matmul"):
    """pressure.py:111-130: periodic x, Neumann y."""
    grid = v[0].grid
    rhs = stencils.divergence(v)
    ops = [laplacian_matrix(grid.shape[0], grid.step[0]), laplacian_matrix_neumann(grid.shape[1], grid.step[1])]
    return pseudoinverse(ops, rhs.dtype, circulant=False, implementation=implementation)(rhs)


def solve_fast_diag_far_field(v):
    """pressure.py:134-154.  Note: the mean-removed rhs is computed (:144) but the solve is applied
    to the *untransformed* rhs (:154); the pseudo-inverse annihilates the constant mode anyway."""
    grid = v[0].grid
    rhs = stencils.divergence(v)
    pbc = pressure_bc_from_velocity(v)
    ops = laplacian_matrix_w_boundaries(grid, grid.cell_center, pbc)
    return pseudoinverse(ops, rhs.dtype, circulant=False, implementation="matmul")(rhs)


SOLVERS = {
    "periodic": solve_fast_diag,
    "channel": solve_fast_diag_moving_wall,
    "far_field": solve_fast_diag_far_field,
}


def projection_and_update_pressure(velocity, pressure: Var, solve="periodic"):
    """pressure.py:58-91: q = solve(v); p_new = q + p_old; v -= forward_difference(q);
    non-periodic: re-impose wall values."""
    grid = velocity[0].grid
    pbc = pressure_bc_from_velocity(velocity)
    solver = SOLVERS[solve] if isinstance(solve, str) else solve
    qsol = solver(velocity)
    q = Var(qsol, grid.cell_center, grid, pbc)
    new_p = Var(qsol + pressure.data, grid.cell_center, grid, pbc)
    periodic = all_periodic(*velocity)
    new_v = []
    for axis, u in enumerate(velocity):
        w = u.with_data(u.data - stencils.forward_difference(q, axis))
        new_v.append(w if periodic else w.impose_bc())
    return tuple(new_v), new_p
