"""ctypes binding of include/jaxib_b200.h.  There is NO CPU fallback: if the native library is
missing and cannot be built, importing the product API raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build


class JaxibError(RuntimeError):
    pass


class Grid(C.Structure):
    """struct jaxib_grid."""

    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("batch", C.c_int32),
        ("lower_x", C.c_double), ("lower_y", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double),
        ("dt", C.c_double), ("nu", C.c_double), ("density", C.c_double),
        ("bc", C.c_int32), ("convect", C.c_int32),
        ("u_wall", C.c_double * 2), ("v_wall", C.c_double * 2), ("force", C.c_double * 2),
        ("ib_window", C.c_int32), ("row0", C.c_int32), ("own_lo", C.c_int32), ("own_hi", C.c_int32),
        ("nx_global", C.c_int32),
        ("u_wall_x", C.c_double * 2), ("v_wall_x", C.c_double * 2),
    ]


class Bodies(C.Structure):
    """struct jaxib_bodies."""

    _fields_ = [
        ("n_bodies", C.c_int32), ("n_markers", C.c_int32),
        ("body_offset", C.c_void_p), ("x0", C.c_void_p), ("y0", C.c_void_p),
        ("max_radius", C.c_double),
        ("marker_pack", C.c_void_p),
        ("tile_flags", C.c_void_p), ("tile_flags_len", C.c_int32),
    ]


MAX_PEERS = 16


class Slab(C.Structure):
    """struct jaxib_slab."""

    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("halo", C.c_int32), ("cpr", C.c_int32),
                ("seam_idle", C.c_int32), ("body_first", C.c_int32), ("body_count", C.c_int32),
                ("marker_first", C.c_int32), ("marker_count", C.c_int32), ("reserved_", C.c_int32)] + [
        (name, C.c_void_p * MAX_PEERS) for name in ("u", "v", "p", "markers", "colbuf", "specbuf", "flags")]


BC = {"periodic": 0, "channel": 1, "far_field": 2}
CONVECT = {"upwind": 0, "van_leer": 1, "van_leer_limiters": 2, "none": 3}
BODY_STATE_STRIDE = 8

_P = C.c_void_p
_SIGNATURES = {
    "jaxib_version": (C.c_char_p, []),
    "jaxib_last_error": (C.c_char_p, []),
    "jaxib_compiled_arch": (C.c_int, []),
    "jaxib_plan_create": (C.c_int, [C.POINTER(Grid), C.POINTER(_P)]),
    "jaxib_plan_destroy": (C.c_int, [_P]),
    "jaxib_scratch_bytes": (C.c_size_t, [_P, C.c_int32]),
    "jaxib_explicit_terms": (C.c_int, [_P, C.POINTER(Grid), _P, _P, _P, _P, _P]),
    "jaxib_explicit_predict": (C.c_int, [_P, C.POINTER(Grid), _P, _P, _P, _P, _P, _P]),
    "jaxib_ib_interp": (C.c_int, [_P, C.POINTER(Grid), C.c_double, C.c_double, _P, _P, _P, C.c_int32, _P, _P, _P]),
    "jaxib_ib_spread_scratch_bytes": (C.c_size_t, [C.POINTER(Grid), C.c_int32]),
    "jaxib_ib_spread": (C.c_int, [_P, C.POINTER(Grid), C.c_double, C.c_double, _P, _P, _P, _P, C.c_int32,
                                  C.c_double, _P, _P, C.c_size_t]),
    "jaxib_ib_force": (C.c_int, [_P, C.POINTER(Grid), C.POINTER(Bodies), _P, _P, _P, _P, C.c_size_t]),
    "jaxib_ib_marker_force": (C.c_int, [_P, C.POINTER(Grid), C.POINTER(Bodies), _P, _P, _P, _P]),
    "jaxib_ib_spread_bodies": (C.c_int, [_P, C.POINTER(Grid), C.POINTER(Bodies), _P, _P, _P, _P]),
    "jaxib_poisson_rows_fwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32]),
    "jaxib_poisson_cols": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32]),
    "jaxib_poisson_rows_inv": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "jaxib_slab_step": (C.c_int, [_P, _P, _P, C.POINTER(Grid), C.POINTER(Slab), C.POINTER(Bodies), _P, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint32), _P, _P]),
    "jaxib_pressure_solve": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_size_t]),
    "jaxib_project": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t]),
    "jaxib_step": (C.c_int, [_P, _P, C.POINTER(Bodies), _P, _P, _P, C.POINTER(C.c_double), C.c_int32, C.c_int32, _P, _P,
                             _P, _P, C.c_size_t]),
    "jaxib_step_profile": (C.c_int, [_P, _P, C.POINTER(Bodies), _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P,
                                     C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_double)]),
    "jaxib_tracer_step": (C.c_int, [_P, C.POINTER(Grid), _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double,
                                    C.c_int32, _P]),
    "jaxib_bilinear_sample": (C.c_int, [_P, C.POINTER(Grid), C.c_double, C.c_double, _P, _P, C.c_int32, _P]),
    "jaxib_pair_force": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double, _P]),
    "jaxib_penalty_perm": (C.c_int, [_P, C.POINTER(Grid), _P, _P, C.c_int32, C.c_int32, C.c_double, C.c_double, _P]),
}

_lock = threading.Lock()
_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib_path() -> str:
    return _build.LIB


def load():
    """Loads (building first if the in-tree .so is absent or stale).  Raises if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        # build() is a no-op when the in-tree .so matches the digest of the sources; it rebuilds a stale one
        # (edited .cu files) under a file lock, so concurrent ranks do not race on the in-tree build
        path = _build.build(force=os.environ.get("JAXIB_REBUILD") == "1")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = header / library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().jaxib_last_error().decode()
        raise JaxibError(f"{what or 'jaxib'} failed (status {rc}): {msg}")
