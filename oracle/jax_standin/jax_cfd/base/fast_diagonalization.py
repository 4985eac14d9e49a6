"""RESTATEMENT of jax_cfd.base.fast_diagonalization (the published algorithm, not the file): the only jax_cfd code the
reference's pressure solve runs (pressure.py:105-107,126-128,149-151).

  pseudoinverse(operators, dtype, hermitian, circulant, implementation, cutoff)
      = transform(v -> 1/v where |v| > cutoff else 0, ...),   cutoff = 10 * eps(dtype)
  transform: with one symmetric operator per axis, A = sum_k I x .. x A_k x .. x I is diagonal in the product of the
      1-D eigenbases, with eigenvalue sum_k lambda_k[i_k]:
        'matmul'  eigh of every A_k (float64, NumPy), eigenvalues outer-summed, func applied, cast to `dtype`;
                  rhs is rotated into the eigenbasis axis by axis (tensordot over the leading axis, which cycles the
                  axes once round), scaled, rotated back
        'rfft'    circulant operators: eigenvalues = fft(first column); rhs -> irfftn(diag * rfftn(rhs)); falls back
                  to 'matmul' when the last axis is odd
      implementation None picks 'rfft' for circulant hermitian operators, else 'matmul'.
Eigen-decompositions are float64 and cast once; the transforms of rhs run in `dtype` (float32 here)."""
import functools

import numpy as _np

from jax.numpy import _plain, _wrap


def transform(func, operators, dtype, *, hermitian=False, circulant=False, implementation=None, precision=None):
    operators = [_np.asarray(_plain(op), dtype=_np.float64) for op in operators]
    for op in operators:
        if op.ndim != 2 or op.shape[0] != op.shape[1]:
            raise ValueError("operators must be square matrices")
    if implementation is None:
        implementation = ("rfft" if hermitian else "fft") if circulant else "matmul"
    if implementation == "rfft" and operators[-1].shape[0] % 2:
        implementation = "matmul"
    dtype = _np.dtype(dtype)
    if implementation == "matmul":
        if not hermitian:
            raise ValueError("matmul needs hermitian operators")
        lam, vec = zip(*(_np.linalg.eigh(op) for op in operators))
        diag = _np.asarray(func(functools.reduce(_np.add.outer, lam)), dtype=dtype)
        vec = [v.astype(dtype) for v in vec]

        def apply(rhs):
            out = _np.asarray(_plain(rhs), dtype=dtype)
            for v in vec:
                out = _np.tensordot(out, v, axes=(0, 0))
            out = out * diag
            for v in vec:
                out = _np.tensordot(out, v, axes=(0, 1))
            return _wrap(out.astype(dtype))
        return apply
    if implementation in ("rfft", "fft"):
        if not circulant:
            raise ValueError("fft needs circulant operators")
        lam = [_np.fft.fft(op[:, 0]) for op in operators]
        summed = functools.reduce(_np.add.outer, lam)
        if implementation == "rfft":
            summed = summed[..., : operators[-1].shape[0] // 2 + 1]
            diag = _np.asarray(func(summed)).astype(_np.complex64 if dtype == _np.float32 else _np.complex128)

            def apply(rhs):
                r = _np.asarray(_plain(rhs), dtype=dtype)
                return _wrap(_np.fft.irfftn(diag * _np.fft.rfftn(r), s=r.shape).astype(dtype))
            return apply
        diag = _np.asarray(func(summed))

        def apply(rhs):
            r = _np.asarray(_plain(rhs))
            return _wrap(_np.fft.ifftn(diag * _np.fft.fftn(r)).astype(r.dtype))
        return apply
    raise NotImplementedError(implementation)


def pseudoinverse(operators, dtype, *, hermitian=False, circulant=False, implementation=None, precision=None, cutoff=None):
    if cutoff is None:
        cutoff = 10 * _np.finfo(_np.dtype(dtype)).eps

    def func(v):
        with _np.errstate(divide="ignore", invalid="ignore"):
            return _np.where(abs(v) > cutoff, 1 / v, 0)
    return transform(func, operators, dtype, hermitian=hermitian, circulant=circulant, implementation=implementation,
                     precision=precision)
