"""Tracer (Brownian) coupling mirrored from the reference's jax_ib/MD package.  The pieces of the
un-vendored jax_md the notebooks touch (space.periodic, quantity.force) live in `MD.space` /
`MD.quantity`."""
from . import CFD_MD_coupling, interaction_potential, quantity, simulate, space  # noqa: F401
