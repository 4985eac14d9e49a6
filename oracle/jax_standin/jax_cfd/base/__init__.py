from . import array_utils, equations, fast_diagonalization, funcutils, pressure, time_stepping  # noqa: F401
