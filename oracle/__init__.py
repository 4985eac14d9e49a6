"""CPU oracle for the jax_ib per-timestep hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a NumPy (float32 by default) restatement of the arithmetic that
hashimmg/jax_IB performs in one `step_fn(All_Variables)`: staggered-grid
advection/diffusion, Gaussian immersed-boundary interpolation and spreading,
fast-diagonalisation Poisson projection, clock/body update, tracer coupling and
the Brinkman penalty mask.  Each function cites the reference file:line it
follows.

PINNED AGAINST THE REFERENCE'S OWN SOURCE FILES, run in the build container.  The
reference ships no tests, golden vectors or fixtures for this path
and JAX is not installable here, so oracle/jax_standin/ provides a
NumPy-backed stand-in for the `jax` entry points the reference uses (plus
tree_math, and containers of jax_md / jax_cfd): with it on sys.path,
tests/golden/make_reference_golden.py imports the UNMODIFIED files under
/root/reference/jax_ib/{base,MD,penalty}/ and executes them on seeded inputs;
tests/test_reference_pinned.py feeds the same inputs to this package and compares
(bit for bit for the array arithmetic; 1e-6-level bars through transcendental
functions and the assembled step).  Covered: SURVEY 8a rows a1-a19, a21, a22,
including the assembled step built exactly as Flapping_Demo.ipynb builds it.
NOT covered, and therefore still anchored only on library-independent identities
(tests/test_oracle_identities.py): the arithmetic INSIDE the absent third-party
packages -- jax_cfd's eigen-solve (restated on both sides), XLA's FFT and
transcendental functions, jax_md's pair potentials and JAX's threefry generator
(kT > 0 takes caller-supplied deviates).

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
