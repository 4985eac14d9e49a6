"""jax.interpreters.xla stand-in: DeviceArray is only used in a type annotation of finite_differences.py."""
from ..numpy import Arr as DeviceArray  # noqa: F401
