import dataclasses as _dc

dataclass = _dc.dataclass


def astuple(x):
    return tuple(getattr(x, f.name) for f in _dc.fields(x))
