"""Placeholder: the reference's equations.py imports jax_cfd's equations module and never uses it on the IB path."""
