"""CPU oracle for the jax_ib per-timestep hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a NumPy (float32 by default) restatement of the arithmetic that
hashimmg/jax_IB performs in one `step_fn(All_Variables)`: staggered-grid
advection/diffusion, Gaussian immersed-boundary interpolation and spreading,
fast-diagonalisation Poisson projection, clock/body update, tracer coupling and
the Brinkman penalty mask.  Each function cites the reference file:line it
follows.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path and cannot be executed in this image (no jax/jaxlib/
jax_cfd/jax_md).  The oracle is therefore pinned only by
library-independent identities (partition of unity, div∘grad = Laplacian so
solve(div grad q) = q, closed-form Laplacian eigenvalues, dense-vs-windowed
Gaussian sums, Taylor-Green decay, Poiseuille steady state) in
tests/test_oracle_*.py, and by golden trajectories it generated itself
(tests/golden/, script tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (jax_ib_b200/) never does.

Third-party arithmetic restated here because it is not under /root/reference:
  * jax_cfd.base.fast_diagonalization.{transform,pseudoinverse} (PyPI jax-cfd,
    unpinned in the reference's setup.py:5)            -> oracle/poisson.py
  * jax.scipy.ndimage.map_coordinates(order=1, mode='constant')
                                                        -> oracle/tracers.py
  * jax_md.simulate.brownian / smap.pair / quantity.force (PyPI jax-md, unpinned)
                                                        -> oracle/tracers.py
  * jax_cfd.base.funcutils.{repeated,trajectory}        -> oracle/stepper.py
"""

from . import field, stencils, ib, poisson, stepper, tracers, penalty, kinematics  # noqa: F401
