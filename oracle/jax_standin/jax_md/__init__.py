"""jax_md STAND-IN (jax-md is a dependency of the reference's MD/ files that is absent here).  TEST INFRASTRUCTURE ONLY.

Just enough for the reference's jax_ib/MD/simulate.py and interaction_potential.py to IMPORT and for
brownian_cfd's init_fn / apply_fn (simulate.py:41-99) to run with a caller-supplied force function and shift: the
BrownianState container, canonicalize_force for a function that already returns forces, static_cast, free-space
shift.  Pair potentials (smap.pair), neighbour lists and energy minimisation are NOT restated here; the oracle's
harmonic-Morse pair force is anchored on a finite-difference identity instead (tests/test_oracle_identities.py)."""
from . import dataclasses, energy, minimize, partition, quantity, simulate, smap, space, util  # noqa: F401
