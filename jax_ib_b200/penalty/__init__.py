"""Brinkman penalisation mirrored from the reference's jax_ib/penalty package."""
from . import parameteric_fns, util_funs  # noqa: F401
