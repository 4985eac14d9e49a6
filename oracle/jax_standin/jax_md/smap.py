"""Placeholder for jax_md.smap: `pair` hands back a marker object; the pair potential is NOT restated (see
jax_md/__init__.py).  quantity.force of that marker is a zero force, which is only acceptable where the tracers are
passive and their positions are not compared (the journal-bearing notebook's FLOW fields)."""


class UnrestatedPairEnergy:
    def __init__(self, fn, displacement, kwargs):
        self.fn, self.displacement, self.kwargs = fn, displacement, kwargs


def pair(fn, displacement_or_metric, species=None, **kwargs):
    return UnrestatedPairEnergy(fn, displacement_or_metric, dict(kwargs, species=species))
