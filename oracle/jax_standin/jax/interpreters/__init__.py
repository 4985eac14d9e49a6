from . import xla  # noqa: F401
