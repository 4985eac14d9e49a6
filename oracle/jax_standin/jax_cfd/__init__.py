"""jax_cfd STAND-IN: google/jax-cfd is a dependency of the reference that is absent here (setup.py names it without a
version; the reference's files are forks of jax-cfd 0.2.x modules).  TEST INFRASTRUCTURE ONLY.  Only
`base.fast_diagonalization` does any work -- a restatement of the published algorithm the reference's pressure.py
calls; the other modules exist so that `from jax_cfd.base import ...` lines at the top of the reference's files
resolve (default arguments that are always overridden on the IB path)."""
from . import base  # noqa: F401
