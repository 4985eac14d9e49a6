from . import ndimage, sparse  # noqa: F401
