"""Placeholder for the DEFAULT arguments `solve=pressure.solve_fast_diag` in the reference (pressure.py:58,
equations.py:201,246); every call site on the IB path passes the reference's own solver instead."""


def solve_fast_diag(*a, **k):
    raise NotImplementedError("jax_cfd's own solver is not part of the pinned path; pass jax_ib.base.pressure.solve_*")
