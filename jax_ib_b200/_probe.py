"""Build-time plug-in recognition.

The reference's solver is assembled from user callables (equations.py:194-208), usually small lambdas
around library functions (Flapping_Demo.ipynb cell 8).  JAX traces them; here the assembly step calls
each plug-in ONCE with probe objects: this package's own library functions recognise a probe and
return a tag instead of computing, so `semi_implicit_navier_stokes_timeBC` learns which fused CUDA
path the user's lambdas describe.  A plug-in that raises on a probe, or returns anything but tags,
is treated as opaque user code and executed stage by stage on real data every step.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Optional


class Probe:
    """Stands in for a GridVariable / array / All_Variables during recognition."""

    def __init__(self, kind="array", index=None, parent=None):
        self.kind, self.index, self.parent = kind, index, parent

    # All_Variables-like attribute access
    @property
    def velocity(self):
        return (Probe("velocity", 0, self), Probe("velocity", 1, self))

    @property
    def particles(self):
        return Probe("particles", None, self)

    @property
    def pressure(self):
        return Probe("pressure", None, self)

    def __iter__(self):
        if self.kind == "velocity_vector":
            return iter((Probe("velocity", 0, self), Probe("velocity", 1, self)))
        raise TypeError("probe is not iterable")

    def __len__(self):
        if self.kind == "velocity_vector":
            return 2
        raise TypeError("probe has no len")


def velocity_probe():
    return Probe("velocity_vector")


def is_probe(x) -> bool:
    return isinstance(x, Probe)


@dataclasses.dataclass(frozen=True)
class Tag:
    what: str
    value: Any = None
    extra: Any = None


def tags_of(result, what: str):
    """Returns the list of tag values if `result` is a tag / tuple of tags of kind `what`, else None."""
    items = result if isinstance(result, (tuple, list)) else (result,)
    if not items or not all(isinstance(t, Tag) and t.what == what for t in items):
        return None
    return [t for t in items]


def try_probe(fn, *args):
    """Calls a user plug-in with probes; any failure means 'opaque'."""
    try:
        return fn(*args)
    except Exception:
        return None
