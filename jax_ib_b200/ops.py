"""Stage-level Python API over the C ABI (torch CUDA tensors own the device memory; every call
enqueues on torch's current stream and never synchronises).

These functions are what the reference-shaped plug-ins in this package (advection.py, pressure.py,
IBM_Force.py, ...) and the fused `equations.semi_implicit_navier_stokes_timeBC` are built from.
The product path never touches oracle/ and has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BC, CONVECT, BODY_STATE_STRIDE


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.JaxibError("expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise _lib.JaxibError("expected a contiguous tensor")
    return t.data_ptr()


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise _lib.JaxibError(f"expected float32, got {t.dtype}")
    return t


@dataclasses.dataclass(frozen=True)
class GridSpec:
    """Mirror of struct jaxib_grid (Python floats stay doubles; see the header)."""

    nx: int
    ny: int
    lower: Tuple[float, float] = (0.0, 0.0)
    step: Tuple[float, float] = (1.0, 1.0)
    dt: float = 0.0
    nu: Optional[float] = None  # viscosity / density
    density: float = 1.0
    bc: str = "periodic"
    convect: str = "upwind"
    u_wall: Tuple[float, float] = (0.0, 0.0)
    v_wall: Tuple[float, float] = (0.0, 0.0)
    force: Tuple[float, float] = (0.0, 0.0)
    ib_window: int = 6
    batch: int = 1
    # slab decomposition (see the header): global row of local row 0, marker ownership range, global nx
    row0: int = 0
    own_lo: int = 0
    own_hi: int = 0
    nx_global: int = 0
    # far field (Dirichlet on all walls): velocity values on the x walls (u_wall / v_wall are the y walls)
    u_wall_x: Tuple[float, float] = (0.0, 0.0)
    v_wall_x: Tuple[float, float] = (0.0, 0.0)

    def c_struct(self) -> _lib.Grid:
        g = _lib.Grid()
        g.nx, g.ny, g.batch = int(self.nx), int(self.ny), int(self.batch)
        g.lower_x, g.lower_y = float(self.lower[0]), float(self.lower[1])
        g.dx, g.dy = float(self.step[0]), float(self.step[1])
        g.dt = float(self.dt)
        g.nu = -1.0 if self.nu is None else float(self.nu)
        g.density = float(self.density)
        g.bc = BC[self.bc]
        g.convect = CONVECT[self.convect]
        g.u_wall[0], g.u_wall[1] = (float(x) for x in self.u_wall)
        g.v_wall[0], g.v_wall[1] = (float(x) for x in self.v_wall)
        g.force[0], g.force[1] = (float(x) for x in self.force)
        g.ib_window = int(self.ib_window)
        g.row0, g.own_lo, g.own_hi, g.nx_global = int(self.row0), int(self.own_lo), int(self.own_hi), int(self.nx_global)
        g.u_wall_x[0], g.u_wall_x[1] = (float(x) for x in self.u_wall_x)
        g.v_wall_x[0], g.v_wall_x[1] = (float(x) for x in self.v_wall_x)
        return g

    @property
    def shape(self):
        return (self.nx, self.ny) if self.batch == 1 else (self.batch, self.nx, self.ny)

    def replace(self, **kw) -> "GridSpec":
        return dataclasses.replace(self, **kw)


def _check_field(spec: GridSpec, t: torch.Tensor, name: str):
    _f32(t)
    if t.numel() != spec.batch * spec.nx * spec.ny:
        raise _lib.JaxibError(f"{name}: {tuple(t.shape)} does not match grid {spec.batch}x{spec.nx}x{spec.ny}")


# ------------------------------------------------------------------------------------------------
def explicit_terms(spec: GridSpec, u, v, perm=None):
    """equations.py:42-83: returns (du/dt, dv/dt) = convect + nu*laplacian + force [- perm*v]."""
    _check_field(spec, u, "u"); _check_field(spec, v, "v")
    du, dv = torch.empty_like(u), torch.empty_like(v)
    g = spec.c_struct()
    _lib.check(_lib.load().jaxib_explicit_terms(_stream(), C.byref(g), _ptr(u), _ptr(v), _ptr(perm), _ptr(du), _ptr(dv)),
               "explicit_terms")
    return du, dv


def explicit_predict(spec: GridSpec, u, v, p=None, perm=None, out=None):
    """time_stepping.py:191-208: (u*, v*) = (u, v) + dt*F(u, v) - grad(p)."""
    _check_field(spec, u, "u"); _check_field(spec, v, "v")
    us, vs = (torch.empty_like(u), torch.empty_like(v)) if out is None else out
    g = spec.c_struct()
    _lib.check(_lib.load().jaxib_explicit_predict(_stream(), C.byref(g), _ptr(u), _ptr(v), _ptr(p), _ptr(perm),
                                                  _ptr(us), _ptr(vs)), "explicit_predict")
    return us, vs


def ib_interp(spec: GridSpec, field, offset, xp, yp, return_cells=False):
    """convolution_functions.py:11-37 (windowed): U_k at the n points; optionally the owning cells."""
    _check_field(spec.replace(batch=1), field, "field")
    n = xp.numel()
    out = torch.empty(n, dtype=torch.float32, device=field.device)
    ci = cj = None
    if return_cells:
        ci = torch.empty(n, dtype=torch.int32, device=field.device)
        cj = torch.empty(n, dtype=torch.int32, device=field.device)
    g = spec.replace(batch=1).c_struct()
    _lib.check(_lib.load().jaxib_ib_interp(_stream(), C.byref(g), float(offset[0]), float(offset[1]), _ptr(field),
                                           _ptr(_f32(xp)), _ptr(_f32(yp)), n, _ptr(out), _ptr(ci), _ptr(cj)), "ib_interp")
    return (out, ci, cj) if return_cells else out


def ib_spread(spec: GridSpec, offset, F, xp, yp, dS, out=None, scale=1.0, binned=True):
    """IBM_Force.py:66-87 (windowed): out += scale * sum_k F_k delta(r_k) dS_k.  binned: scratch for the cell-binned
    gather (markers counting-sorted by cell); False: 16 bytes, every tile of the bounding box culls all points."""
    if out is None:
        out = torch.zeros(spec.nx, spec.ny, dtype=torch.float32, device=F.device)
    g = spec.replace(batch=1).c_struct()
    lib = _lib.load()
    nbytes = int(lib.jaxib_ib_spread_scratch_bytes(C.byref(g), F.numel())) if binned else 16
    scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=F.device)
    _lib.check(lib.jaxib_ib_spread(_stream(), C.byref(g), float(offset[0]), float(offset[1]), _ptr(_f32(F)),
                                   _ptr(_f32(xp)), _ptr(_f32(yp)), _ptr(_f32(dS)), F.numel(), float(scale),
                                   _ptr(out), _ptr(scratch), scratch.numel()), "ib_spread")
    return out


class DeviceBodies:
    """struct jaxib_bodies with its device arrays kept alive."""

    def __init__(self, x0_list: Sequence[np.ndarray], y0_list: Sequence[np.ndarray], device):
        sizes = [len(x) for x in x0_list]
        self.n_bodies = len(sizes)
        self.n_markers = int(sum(sizes))
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        x0 = np.concatenate([np.asarray(x, dtype=np.float32) for x in x0_list]) if sizes else np.zeros(0, np.float32)
        y0 = np.concatenate([np.asarray(y, dtype=np.float32) for y in y0_list]) if sizes else np.zeros(0, np.float32)
        self.max_radius = float(np.sqrt(x0.astype(np.float64) ** 2 + y0.astype(np.float64) ** 2).max()) if sizes else 0.0
        self.offset_host = off  # marker range of every body (host copy: slab hints)
        self.body_offset = torch.from_numpy(off).to(device)
        self.x0 = torch.from_numpy(x0).to(device)
        self.y0 = torch.from_numpy(y0).to(device)
        self.offsets_host = off
        # per-marker pack (see struct jaxib_bodies): next marker on the closed curve + owning body
        pack = np.zeros((self.n_markers, 4), dtype=np.float32)
        body = np.zeros(self.n_markers, dtype=np.int32)
        for b in range(self.n_bodies):
            k0, k1 = int(off[b]), int(off[b + 1])
            nxt = np.roll(np.arange(k0, k1), -1)
            pack[k0:k1, 0], pack[k0:k1, 1] = x0[nxt], y0[nxt]
            body[k0:k1] = b
        pack[:, 2] = body.view(np.float32)
        self.marker_pack = torch.from_numpy(pack).to(device)

    def tile_flags(self, spec: "GridSpec") -> torch.Tensor:
        """Zeroed scratch for the occupied-tile marks of struct jaxib_bodies (one per stepping grid)."""
        hmin = min(spec.step)
        he = int(math.ceil(self.max_radius / hmin)) + int(spec.ib_window) + 2
        tiles = (2 * he + 1 + 7) // 8
        return torch.zeros(max(1, spec.batch * self.n_bodies * 2 * tiles * tiles), dtype=torch.int32, device=self.x0.device)

    def c_struct(self, flags: Optional[torch.Tensor] = None) -> _lib.Bodies:
        b = _lib.Bodies()
        b.n_bodies, b.n_markers = self.n_bodies, self.n_markers
        b.body_offset, b.x0, b.y0 = self.body_offset.data_ptr(), self.x0.data_ptr(), self.y0.data_ptr()
        b.max_radius = self.max_radius
        b.marker_pack = self.marker_pack.data_ptr() if self.n_markers else None
        if flags is not None:
            b.tile_flags, b.tile_flags_len = flags.data_ptr(), flags.numel()
        return b


def ib_force(spec: GridSpec, bodies: DeviceBodies, body_state, u_star, v_star):
    """IBM_Force.py:109-115 + time_stepping.py:223, in place: (u*, v*) += dt * f_IB.
    Returns the marker scratch view [5, batch, n_markers] = xp, yp, dS, Fx, Fy."""
    _check_field(spec, u_star, "u_star"); _check_field(spec, v_star, "v_star")
    nbytes = 5 * spec.batch * bodies.n_markers * 4 + 512
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=u_star.device)
    g, b = spec.c_struct(), bodies.c_struct()
    _lib.check(_lib.load().jaxib_ib_force(_stream(), C.byref(g), C.byref(b), _ptr(_f32(body_state)), _ptr(u_star),
                                          _ptr(v_star), _ptr(scratch), nbytes), "ib_force")
    base = (scratch.data_ptr() + 255) & ~255
    skip = base - scratch.data_ptr()
    return scratch[skip:skip + 5 * spec.batch * bodies.n_markers * 4].view(torch.float32).view(5, spec.batch, bodies.n_markers)


def ib_marker_force(spec: GridSpec, bodies: DeviceBodies, body_state, u_star, v_star, markers):
    """First half of ib_force: fills markers [5, batch, n_markers] = xp, yp, dS, Fx, Fy (F = 0 for markers
    outside spec.own_lo .. own_hi when that range is set)."""
    _check_field(spec, u_star, "u_star"); _check_field(spec, v_star, "v_star")
    if markers.numel() < 5 * spec.batch * bodies.n_markers:
        raise _lib.JaxibError("markers buffer too small")
    g, b = spec.c_struct(), bodies.c_struct()
    _lib.check(_lib.load().jaxib_ib_marker_force(_stream(), C.byref(g), C.byref(b), _ptr(_f32(body_state)), _ptr(u_star),
                                                 _ptr(v_star), _ptr(_f32(markers))), "ib_marker_force")
    return markers


def ib_spread_bodies(spec: GridSpec, bodies: DeviceBodies, body_state, markers, u_star, v_star):
    """Second half of ib_force: (u*, v*) += dt * spread(markers), in place."""
    _check_field(spec, u_star, "u_star"); _check_field(spec, v_star, "v_star")
    g, b = spec.c_struct(), bodies.c_struct()
    _lib.check(_lib.load().jaxib_ib_spread_bodies(_stream(), C.byref(g), C.byref(b), _ptr(_f32(body_state)),
                                                  _ptr(_f32(markers)), _ptr(u_star), _ptr(v_star)), "ib_spread_bodies")


class Plan:
    """jaxib_plan + a scratch buffer sized for the fused step."""

    def __init__(self, spec: GridSpec, n_markers: int = 0, device="cuda", alloc_scratch: bool = True):
        self.spec = spec
        self.device = torch.device(device)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        g = spec.c_struct()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.jaxib_plan_create(C.byref(g), C.byref(self._h)), "plan_create")
        self.n_markers = n_markers
        self.scratch_bytes = int(self._lib.jaxib_scratch_bytes(self._h, n_markers))
        # stage-level users (slab.py) bring their own buffers
        self.scratch = torch.empty(self.scratch_bytes, dtype=torch.uint8, device=self.device) if alloc_scratch else None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.jaxib_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def pressure_solve(self, u, v):
        """pressure.py:94-130: q = pinv(L)(div(u, v))."""
        _check_field(self.spec, u, "u"); _check_field(self.spec, v, "v")
        q = torch.empty_like(u)
        _lib.check(self._lib.jaxib_pressure_solve(_stream(), self._h, _ptr(u), _ptr(v), _ptr(q), _ptr(self.scratch),
                                                  self.scratch_bytes), "pressure_solve")
        return q

    def project(self, u, v, p, inplace=False):
        """pressure.py:58-91: returns (u, v, p) projected / updated."""
        for t, n in ((u, "u"), (v, "v"), (p, "p")):
            _check_field(self.spec, t, n)
        uo, vo, po = (u, v, p) if inplace else (torch.empty_like(u), torch.empty_like(v), torch.empty_like(p))
        _lib.check(self._lib.jaxib_project(_stream(), self._h, _ptr(u), _ptr(v), _ptr(p), _ptr(uo), _ptr(vo), _ptr(po),
                                           _ptr(self.scratch), self.scratch_bytes), "project")
        return uo, vo, po

    # -- stage-level (slab decomposition: the caller exchanges data between the stages) ----------
    def rows_fwd(self, u, v, spectrum, sp=0):
        """divergence + row FFT of the plan's nx rows starting at u / v (slab mode reads u row -1 in place)."""
        _lib.check(self._lib.jaxib_poisson_rows_fwd(_stream(), self._h, _ptr(u), _ptr(v), _ptr(spectrum), int(sp)),
                   "poisson_rows_fwd")

    def cols(self, spectrum, sp=0, ncols=0, col0=0):
        _lib.check(self._lib.jaxib_poisson_cols(_stream(), self._h, _ptr(spectrum), int(sp), int(ncols), int(col0)),
                   "poisson_cols")

    def rows_inv(self, spectrum, u, v, p, u_out, v_out, p_out, sp=0):
        _lib.check(self._lib.jaxib_poisson_rows_inv(_stream(), self._h, _ptr(spectrum), int(sp), _ptr(u), _ptr(v), _ptr(p),
                                                    _ptr(u_out), _ptr(v_out), _ptr(p_out)), "poisson_rows_inv")

    def step(self, u, v, p, nsteps=1, bodies: Optional[DeviceBodies] = None, body_table=None, perm=None,
             incremental_pressure=True, wall_table=None, body_table_host=None, tile_flags=None):
        """time_stepping.py:144-255, `nsteps` times, in place on (u, v, p).  wall_table: host array
        [nsteps, 8] = (u_lo, u_hi, v_lo, v_hi) on the y walls, then on the x walls, per step (boundaries.py:519-542);
        None = constant walls."""
        for t, n in ((u, "u"), (v, "v"), (p, "p")):
            _check_field(self.spec, t, n)
        b = bodies.c_struct(tile_flags) if bodies is not None else None
        if bodies is not None:
            if bodies.n_markers > self.n_markers:
                raise _lib.JaxibError("plan scratch was sized for fewer markers")
            need = nsteps * self.spec.batch * bodies.n_bodies * BODY_STATE_STRIDE
            if body_table is None or body_table.numel() < need:
                raise _lib.JaxibError("body_table too small for nsteps")
        walls = None
        if wall_table is not None:
            w = np.ascontiguousarray(np.asarray(wall_table, dtype=np.float64).reshape(-1, 8))
            if w.shape[0] < nsteps:
                raise _lib.JaxibError("wall_table too small for nsteps")
            walls = w.ctypes.data_as(C.POINTER(C.c_double))
        host = None
        if body_table_host is not None and bodies is not None:
            # the same table on the host: lets the library overlap the IB chain with the predictor (see the header)
            th = np.ascontiguousarray(np.asarray(body_table_host, dtype=np.float32))
            if th.size >= nsteps * self.spec.batch * bodies.n_bodies * BODY_STATE_STRIDE:
                host = th.ctypes.data_as(C.c_void_p)
        _lib.check(self._lib.jaxib_step(_stream(), self._h, C.byref(b) if b is not None else None, _ptr(body_table),
                                        host, _ptr(perm), walls, 1 if incremental_pressure else 0, int(nsteps), _ptr(u),
                                        _ptr(v), _ptr(p), _ptr(self.scratch), self.scratch_bytes), "step")
        return u, v, p


STAGES = ("explicit_predict", "ib_interp_force", "ib_spread", "rows_fwd", "cols", "rows_inv", "step")


def step_profile(plan: Plan, u, v, p, nsteps, bodies=None, body_table=None, perm=None, incremental_pressure=True,
                 flush: Optional[torch.Tensor] = None, tile_flags=None):
    """Measurement aid: like Plan.step but returns {stage: summed ms} from per-stage CUDA events.
    `flush` (a device buffer larger than L2) is overwritten before every step, outside the brackets.
    Synchronises the stream."""
    b = bodies.c_struct(tile_flags) if bodies is not None else None
    ms = (C.c_double * len(STAGES))()
    _lib.check(plan._lib.jaxib_step_profile(
        _stream(), plan._h, C.byref(b) if b is not None else None, _ptr(body_table), _ptr(perm),
        1 if incremental_pressure else 0, int(nsteps), _ptr(u), _ptr(v), _ptr(p), _ptr(plan.scratch), plan.scratch_bytes,
        _ptr(flush), flush.numel() * flush.element_size() if flush is not None else 0, ms), "step_profile")
    return dict(zip(STAGES, [float(x) for x in ms]))


# ---- tracers / penalty ---------------------------------------------------------------------------
def bilinear_sample(spec: GridSpec, field, offset, R):
    n = R.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=field.device)
    g = spec.replace(batch=1).c_struct()
    _lib.check(_lib.load().jaxib_bilinear_sample(_stream(), C.byref(g), float(offset[0]), float(offset[1]), _ptr(field),
                                                 _ptr(_f32(R)), n, _ptr(out)), "bilinear_sample")
    return out


def pair_force(R, species, h, r0, box, D0=0.0, alpha=10.0, k=1.0):
    F = torch.empty_like(R)
    ns = h.shape[0]
    _lib.check(_lib.load().jaxib_pair_force(_stream(), _ptr(_f32(R)), _ptr(species), R.shape[0], ns, _ptr(_f32(h)),
                                            _ptr(_f32(r0)), float(D0), float(alpha), float(k), float(box[0]),
                                            float(box[1]), _ptr(F)), "pair_force")
    return F


def tracer_step(spec: GridSpec, u, v, R, dt_md, gamma=1.0, mass=1.0, kT=0.0, pair=None, xi=None, is_mobile=None):
    """MD/simulate.py:81-97 in place on R [n, 2]."""
    g = spec.replace(batch=1).c_struct()
    _lib.check(_lib.load().jaxib_tracer_step(_stream(), C.byref(g), _ptr(u), _ptr(v), _ptr(pair), _ptr(xi),
                                             _ptr(is_mobile), float(dt_md), float(mass * gamma), float(kT), R.shape[0],
                                             _ptr(_f32(R))), "tracer_step")
    return R


def penalty_perm(spec: GridSpec, centers, Rtheta, K, width):
    """penalty/util_funs.py:76-123 with the tanh smoothing: returns perm [nx, ny]."""
    perm = torch.empty(spec.nx, spec.ny, dtype=torch.float32, device=centers.device)
    g = spec.replace(batch=1).c_struct()
    _lib.check(_lib.load().jaxib_penalty_perm(_stream(), C.byref(g), _ptr(_f32(centers)), _ptr(_f32(Rtheta)),
                                              Rtheta.shape[0], Rtheta.shape[1], float(K), float(width), _ptr(perm)),
               "penalty_perm")
    return perm
