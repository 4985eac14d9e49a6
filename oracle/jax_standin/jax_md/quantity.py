import jax.numpy as jnp


def canonicalize_force(energy_or_force_fn):
    """jax-md probes the function and differentiates it when it returns an energy; the generators only ever pass a
    function that already returns forces."""
    return energy_or_force_fn


def force(energy_fn):
    """-grad of an energy needs reverse-mode autodiff, which the stand-in does not have.  For the unrestated pair energy
    of smap.pair the result is a ZERO force (passive tracers without wall repulsion: their positions are then not the
    reference's and must not be compared); anything else is refused."""
    from .smap import UnrestatedPairEnergy
    if isinstance(energy_fn, UnrestatedPairEnergy):
        return lambda R, **kwargs: jnp.zeros_like(R)
    raise NotImplementedError("quantity.force: autodiff is not restated")
