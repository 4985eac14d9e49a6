def canonicalize_force(energy_or_force_fn):
    """jax-md probes the function and differentiates it when it returns an energy; the generator only ever passes a
    function that already returns forces."""
    return energy_or_force_fn
