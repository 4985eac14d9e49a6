"""Slab decomposition of ONE large grid over several GPUs (BASELINE config 4).

New capability: the reference only shards markers with `jax.pmap` and replicates the grid
(IBM_Force.py:77-87, convolution_functions.py:28-35).  Here the x axis (rows; y stays contiguous) is
split into `world` slabs of `nx / world` rows.  Every rank keeps its rows plus `halo` = 2R + 4 rows of
its neighbours on either side, so that one step needs exactly three exchanges:

    K1 predictor on the extended slab (the halo absorbs the stencil; the local array is treated as a
       periodic grid of its own, garbage stays inside the outermost two halo rows)
    K2 forces of the markers whose spreading window touches the slab; their interpolation windows lie
       inside the halo, so every slab computes them itself, bit for bit as one GPU would -- no exchange
    K3 spreading onto the local rows (all markers, culled per tile as on one GPU)
    K4 divergence + row FFT of the owned rows -> spectrum [nloc][world * cpr]
  X1 all-to-all: rank r receives the columns [r * cpr, (r+1) * cpr) of every row
    K5 column FFT x eigenvalues x inverse over all nx rows of the local columns
  X2 all-to-all back; every destination also gets the first row of the NEXT slab (K6 differences q
       forward in x), so no separate halo pass is needed for q
    K6 inverse row FFT + projection + pressure update on the owned rows
  X3 halo exchange of the new u, v, p rows with the two neighbours (periodic ring)

Cells next to the periodic seam: the reference takes no periodic image of a marker (raw xp - X), so
the IB kernels see the local array as un-wrapped segments -- the rows whose global index lies in
[0, nx) -- plus, on the first slab, the single wrapped row nx-1 that K4's backward difference reads.

The per-rank compute (`SlabRank`) is separated from the exchanges (`DistComm`: torch.distributed,
NCCL over NVLink on GPUs, gloo in the CPU tests of the exchange logic; `LocalComm`: several virtual
slabs inside one process on one GPU), so the decomposition is tested against the single-GPU step on
one GPU and the collectives on real multi-GPU boxes.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import kinematics, ops

INT_MAX = 2 ** 31 - 1


@dataclasses.dataclass(frozen=True)
class Segment:
    """Local rows [start, start + rows) hold the global rows [global0, global0 + rows), un-wrapped."""
    start: int
    rows: int
    global0: int


class SlabLayout:
    """Pure index arithmetic of one rank's slab (host only; covered by the CPU tests)."""

    def __init__(self, nx: int, ny: int, world: int, rank: int, halo: int):
        if nx % world:
            raise ValueError(f"nx={nx} is not divisible by the number of slabs {world}")
        if ny % 2:
            raise ValueError("ny must be even")
        self.nx, self.ny, self.world, self.rank, self.halo = nx, ny, world, rank, halo
        self.nloc = nx // world
        if halo > self.nloc:
            raise ValueError(f"halo {halo} exceeds the slab height {self.nloc}")
        self.own_lo = rank * self.nloc
        self.own_hi = self.own_lo + self.nloc
        self.nxl = self.nloc + 2 * halo
        self.row0 = self.own_lo - halo  # global row of local row 0 (may be negative: wrapped)
        n2 = ny // 2
        # columns per rank: the n2 + 1 wavenumbers (+3 padding columns the row kernels write), rounded to 8
        per_rank = -(-(n2 + 4) // world)
        self.cpr = -(-per_rank // 8) * 8
        self.sp = world * self.cpr

    # Markers a slab needs the force of: every marker whose spreading window touches the u** rows
    # [own_lo - 1, own_hi) or the v** rows [own_lo, own_hi), i.e. u-cell row in [own_lo - R - 2, own_hi + R).
    # Their interpolation windows reach 2R + 2 rows past the owned block, so with a halo of 2R + 4 rows
    # (2 more for the predictor's stencil) every slab computes these forces itself: nothing is exchanged.
    # The first / last slab also take the markers that lie off the grid.
    def marker_range(self, R: int):
        lo = -INT_MAX if self.rank == 0 else self.own_lo - R - 2
        hi = INT_MAX if self.rank == self.world - 1 else self.own_hi + R
        return lo, hi

    def top_segment(self, R: int):
        """First slab of several: the seam row nx-1 is fed by markers near the TOP of the domain, whose
        windows lie in the wrapped lower halo.  Returns (Segment, marker_lo, marker_hi) or None."""
        if self.rank == 0 and self.world > 1:
            return Segment(0, self.halo, self.nx - self.halo), self.nx - R - 3, INT_MAX
        return None

    def main_segment(self) -> Segment:
        a = max(0, -self.row0)
        b = min(self.nxl, self.nx - self.row0)
        return Segment(a, b - a, self.row0 + a)

    def seam_row(self) -> Optional[Segment]:
        """The wrapped row below the first owned row (global nx-1 on the first slab), else None."""
        if self.own_lo == 0:
            return Segment(self.halo - 1, 1, self.nx - 1)
        return None

    def inverse_row_index(self) -> np.ndarray:
        """Rows of the [nx][cpr] column buffer sent back to each destination: its nloc rows + the next one."""
        idx = (np.arange(self.world)[:, None] * self.nloc + np.arange(self.nloc + 1)[None, :]) % self.nx
        return idx.reshape(-1).astype(np.int64)

    def global_rows(self) -> np.ndarray:
        """Global (wrapped) row of every local row."""
        return (self.row0 + np.arange(self.nxl)) % self.nx


# ------------------------------------------------------------------------------------------------
# exchanges
class DistComm:
    """torch.distributed exchanges for ONE local rank (lists of length 1)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def transpose(self, send: Sequence[torch.Tensor], recv: Sequence[torch.Tensor]):
        dist.all_to_all_single(recv[0], send[0], group=self.group)

    def halos(self, lo_send, hi_send, lo_recv, hi_recv):
        """lo_send[0] / hi_send[0]: lists of contiguous tensors (first / last owned rows of u, v, p);
        lo_recv / hi_recv: the halo views they land in on the neighbours."""
        prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        if self.group is not None:
            prev, nxt = dist.get_global_rank(self.group, prev), dist.get_global_rank(self.group, nxt)
        ops_ = []
        # message order per peer pair is the same on both sides (first rows up, last rows down)
        for t in lo_send[0]:
            ops_.append(dist.P2POp(dist.isend, t, prev, self.group))
        for t in hi_send[0]:
            ops_.append(dist.P2POp(dist.isend, t, nxt, self.group))
        for t in hi_recv[0]:
            ops_.append(dist.P2POp(dist.irecv, t, nxt, self.group))
        for t in lo_recv[0]:
            ops_.append(dist.P2POp(dist.irecv, t, prev, self.group))
        for w in dist.batch_isend_irecv(ops_):
            w.wait()


class LocalComm:
    """The same exchanges between `world` virtual slabs living in one process (one GPU)."""

    def __init__(self, world: int):
        self.world = world

    def transpose(self, send, recv):
        # send[r][d] goes to recv[d][r]
        for r in range(self.world):
            for d in range(self.world):
                recv[d][r].copy_(send[r][d])

    def halos(self, lo_send, hi_send, lo_recv, hi_recv):
        for r in range(self.world):
            prev, nxt = (r - 1) % self.world, (r + 1) % self.world
            for src, dst in zip(lo_send[r], hi_recv[prev]):
                dst.copy_(src)
            for src, dst in zip(hi_send[r], lo_recv[nxt]):
                dst.copy_(src)


# ------------------------------------------------------------------------------------------------
class SlabRank:
    """Device state and kernels of one slab."""

    def __init__(self, spec: ops.GridSpec, layout: SlabLayout, bodies: Optional[ops.DeviceBodies], device):
        if spec.bc != "periodic":
            raise NotImplementedError("slab decomposition: periodic grids only")
        if spec.batch != 1:
            raise NotImplementedError("slab decomposition: batch = 1")
        self.spec, self.lay, self.bodies, self.device = spec, layout, bodies, torch.device(device)
        L = layout
        f32 = dict(dtype=torch.float32, device=self.device)
        self.u, self.v, self.p = (torch.zeros(L.nxl, L.ny, **f32) for _ in range(3))
        self.us, self.vs = torch.zeros(L.nxl, L.ny, **f32), torch.zeros(L.nxl, L.ny, **f32)
        self.spec_ext = spec.replace(nx=L.nxl, row0=L.row0)
        seg = L.main_segment()
        self.seg = seg
        mlo, mhi = L.marker_range(spec.ib_window)
        self.spec_ib = spec.replace(nx=seg.rows, row0=seg.global0, own_lo=mlo, own_hi=mhi)
        top = L.top_segment(spec.ib_window)
        self.top = None if top is None else top[0]
        self.spec_top = None if top is None else spec.replace(nx=top[0].rows, row0=top[0].global0, own_lo=top[1],
                                                              own_hi=top[2])
        seam = L.seam_row()
        self.seam = seam
        self.spec_seam = None if seam is None else spec.replace(nx=1, row0=seam.global0)
        self.plan_rows = ops.Plan(spec.replace(nx=L.nloc, nx_global=L.nx), 0, self.device, alloc_scratch=False)
        self.plan_cols = ops.Plan(spec.replace(nx=L.nx), 0, self.device, alloc_scratch=False)
        M = bodies.n_markers if bodies is not None else 0
        self.markers = torch.zeros(5, 1, max(M, 1), **f32)
        # spectrum buffers (complex as trailing [2] float32)
        self.spectrum = torch.zeros(L.nloc + 1, L.world, L.cpr, 2, **f32)
        self.send = torch.zeros(L.world, L.nloc, L.cpr, 2, **f32)
        self.recv = torch.zeros(L.world, L.nloc, L.cpr, 2, **f32)        # == [nx][cpr] complex
        self.send2 = torch.zeros(L.world, L.nloc + 1, L.cpr, 2, **f32)
        self.recv2 = torch.zeros(L.world, L.nloc + 1, L.cpr, 2, **f32)
        self.inv_idx = torch.from_numpy(L.inverse_row_index()).to(self.device)

    # -- data movement helpers ---------------------------------------------------------------------
    def load_global(self, u, v, p):
        """Takes this slab's rows (+ halos) out of full [nx, ny] arrays (any device)."""
        rows = torch.from_numpy(self.lay.global_rows()).to(u.device)
        for dst, src in ((self.u, u), (self.v, v), (self.p, p)):
            dst.copy_(src.index_select(0, rows).to(self.device, dtype=torch.float32))

    def owned(self, t: torch.Tensor) -> torch.Tensor:
        h = self.lay.halo
        return t[h:h + self.lay.nloc]

    def _rows(self, t, seg: Segment):
        return t[seg.start:seg.start + seg.rows]

    # -- phases ------------------------------------------------------------------------------------
    def phase_a(self, body_state: Optional[torch.Tensor]):
        """K1 + K2.  body_state: [1, B, 8] float32 (device) for this step."""
        ops.explicit_predict(self.spec_ext, self.u, self.v, self.p, out=(self.us, self.vs))
        if self.bodies is not None:
            ops.ib_marker_force(self.spec_ib, self.bodies, body_state, self._rows(self.us, self.seg),
                                self._rows(self.vs, self.seg), self.markers)
            if self.top is not None:  # markers near the top of the domain feed the first slab's seam row
                ops.ib_marker_force(self.spec_top, self.bodies, body_state, self._rows(self.us, self.top),
                                    self._rows(self.vs, self.top), self.markers)

    def phase_b(self, body_state: Optional[torch.Tensor]):
        """K3 + K4 + pack."""
        if self.bodies is not None:
            ops.ib_spread_bodies(self.spec_ib, self.bodies, body_state, self.markers, self._rows(self.us, self.seg),
                                 self._rows(self.vs, self.seg))
            if self.seam is not None:
                ops.ib_spread_bodies(self.spec_seam, self.bodies, body_state, self.markers,
                                     self._rows(self.us, self.seam), self._rows(self.vs, self.seam))
        L = self.lay
        self.plan_rows.rows_fwd(self.owned(self.us), self.owned(self.vs), self.spectrum, sp=L.sp)
        self.send.copy_(self.spectrum[:L.nloc].permute(1, 0, 2, 3))

    def phase_c(self):
        """K5 on the local columns + pack for the way back."""
        L = self.lay
        self.plan_cols.cols(self.recv, sp=L.cpr, ncols=L.cpr, col0=L.rank * L.cpr)
        torch.index_select(self.recv.view(L.nx, L.cpr * 2), 0, self.inv_idx,
                           out=self.send2.view(L.world * (L.nloc + 1), L.cpr * 2))

    def phase_d(self):
        """unpack + K6 (writes the owned rows of u, v, p in place)."""
        L = self.lay
        self.spectrum.copy_(self.recv2.permute(1, 0, 2, 3))
        uo, vo, po = self.owned(self.u), self.owned(self.v), self.owned(self.p)
        self.plan_rows.rows_inv(self.spectrum, self.owned(self.us), self.owned(self.vs), po, uo, vo, po, sp=L.sp)

    def halo_views(self):
        h, n = self.lay.halo, self.lay.nloc
        fields = (self.u, self.v, self.p)
        lo_send = [t[h:2 * h] for t in fields]          # first owned rows -> previous slab's upper halo
        hi_send = [t[n:n + h] for t in fields]          # last owned rows  -> next slab's lower halo
        lo_recv = [t[0:h] for t in fields]
        hi_recv = [t[h + n:h + n + h] for t in fields]
        return lo_send, hi_send, lo_recv, hi_recv


def seam_is_idle(xc, max_radius: float, ib_window: int, dx: float, lower: float, nx: int) -> bool:
    """Host hint `jaxib_slab.seam_idle`: True when, at every step of the block (xc: [nsteps, B] body-centre x), every
    body stays clear of both ends of the periodic x-range by its radius plus the spreading / interpolation reach, so
    the first slab has no top markers and no seam row to spread onto."""
    reach = max_radius + (ib_window + 4) * dx
    xc = np.asarray(xc, dtype=np.float64)
    return bool(((xc - reach) > lower).all() and ((xc + reach) < lower + nx * dx).all())


def bodies_near(xc, max_radius: float, ib_window: int, dx: float, lower: float, own_lo: int, own_hi: int, first: bool,
                last: bool):
    """Host hint `jaxib_slab.body_first / body_count`: (first, last + 1) of the bodies that can touch the slab
    of rows [own_lo, own_hi) during the block -- a contiguous superset; None = all bodies.  A body counts when its
    centre track, widened by its radius plus the interpolation and spreading reach (2 R + 8 cells, more than the
    2 R + 4 halo), overlaps the slab's rows; the first / last slab also own everything off the grid."""
    xc = np.asarray(xc, dtype=np.float64) - lower
    if xc.ndim != 2 or xc.shape[1] < 2:
        return None
    reach = max_radius + (2 * ib_window + 8) * dx
    lo, hi = xc.min(axis=0) - reach, xc.max(axis=0) + reach
    a = -np.inf if first else own_lo * dx
    b = np.inf if last else own_hi * dx
    near = np.nonzero((hi >= a) & (lo <= b))[0]
    if near.size == 0:
        return (0, 1)  # nothing comes near: one body keeps the launches well-formed (its tiles leave at once)
    return (int(near.min()), int(near.max()) + 1)


PHASES = ("predict_force", "spread_rows_fwd", "x_transpose_fwd", "cols", "x_transpose_inv", "rows_inv", "x_halos")


def _step(ranks: List[SlabRank], comm, body_state, mark=None):
    """One time step.  `mark(k)` (measurement only) is called after phase k of PHASES has been enqueued."""
    mark = mark or (lambda k: None)
    for r in ranks:
        r.phase_a(body_state)
    mark(0)
    for r in ranks:
        r.phase_b(body_state)
    mark(1)
    comm.transpose([r.send for r in ranks], [r.recv for r in ranks])
    mark(2)
    for r in ranks:
        r.phase_c()
    mark(3)
    comm.transpose([r.send2 for r in ranks], [r.recv2 for r in ranks])
    mark(4)
    for r in ranks:
        r.phase_d()
    mark(5)
    views = [r.halo_views() for r in ranks]
    comm.halos(*[[v[k] for v in views] for k in range(4)])
    mark(6)


# ------------------------------------------------------------------------------------------------
# peer-memory mode: the kernels store into the consuming slab's buffers over NVLink (csrc/slab.cu)
def _align(n: int, a: int = 256) -> int:
    return -(-n // a) * a


class PeerBlock:
    """Byte layout of one slab's symmetric allocation (identical on every slab): the buffers of
    struct jaxib_slab, each 256-byte aligned."""

    FLAG_WORDS = 128

    def __init__(self, layout: SlabLayout, n_markers: int):
        L = layout
        sizes = [("u", L.nxl * L.ny * 4), ("v", L.nxl * L.ny * 4), ("p", L.nxl * L.ny * 4),
                 ("markers", 5 * max(n_markers, 1) * 4), ("colbuf", L.nx * L.cpr * 8),
                 ("specbuf", (L.nloc + 1) * L.sp * 8), ("flags", self.FLAG_WORDS * 4)]
        self.offset, off = {}, 0
        for name, nbytes in sizes:
            self.offset[name] = off
            off += _align(nbytes)
        self.nbytes = off
        self.size = dict(sizes)


class PeerRank:
    """One slab in peer-memory mode: views into its own block + the jaxib_slab pointer table."""

    def __init__(self, spec: ops.GridSpec, layout: SlabLayout, bodies, device, block: PeerBlock, local: torch.Tensor,
                 base_ptrs: Sequence[int]):
        import ctypes as C

        from . import _lib
        if spec.bc != "periodic" or spec.batch != 1:
            raise NotImplementedError("slab decomposition: periodic grids, batch = 1")
        self.spec, self.lay, self.bodies, self.device = spec, layout, bodies, torch.device(device)
        L = layout
        self.block, self.local = block, local
        M = bodies.n_markers if bodies is not None else 0

        def view(name, shape):
            o = block.offset[name]
            n = int(np.prod(shape)) * 4
            return local[o:o + n].view(torch.float32).view(*shape)

        self.u, self.v, self.p = (view(k, (L.nxl, L.ny)) for k in "uvp")
        self.markers = view("markers", (5, 1, max(M, 1)))
        self.colbuf = view("colbuf", (L.nx, L.cpr, 2))
        self.flags = local[block.offset["flags"]:block.offset["flags"] + 4 * PeerBlock.FLAG_WORDS].view(torch.int32)
        self.us = torch.zeros(L.nxl, L.ny, dtype=torch.float32, device=self.device)
        self.vs = torch.zeros(L.nxl, L.ny, dtype=torch.float32, device=self.device)
        self.plan_rows = ops.Plan(spec.replace(nx=L.nloc, nx_global=L.nx), 0, self.device, alloc_scratch=False)
        self.plan_cols = ops.Plan(spec.replace(nx=L.nx), 0, self.device, alloc_scratch=False)
        sl = _lib.Slab()
        sl.world, sl.rank, sl.halo, sl.cpr = L.world, L.rank, L.halo, L.cpr
        for d, base in enumerate(base_ptrs):
            for name in ("u", "v", "p", "colbuf", "specbuf", "flags"):
                getattr(sl, name)[d] = base + block.offset[name]
            sl.markers[d] = base + block.offset["markers"] if M else None
        self.c_slab = sl
        self.c_grid = spec.c_struct()  # the GLOBAL grid
        self.epoch = C.c_uint32(0)
        self._lib = _lib

    load_global = SlabRank.load_global
    owned = SlabRank.owned

    def run(self, table, nsteps: int, phase_lo: int = 0, phase_hi: int = 3, barriers: bool = True, seam_idle: bool = False,
            body_range=None):
        import ctypes as C
        self.c_slab.seam_idle = 1 if seam_idle else 0
        b0, b1 = body_range if body_range is not None else (0, 0)
        self.c_slab.body_first, self.c_slab.body_count = int(b0), int(b1 - b0)
        if b1 > b0:
            off = self.bodies.offset_host
            self.c_slab.marker_first, self.c_slab.marker_count = int(off[b0]), int(off[b1] - off[b0])
        else:
            self.c_slab.marker_first = self.c_slab.marker_count = 0
        if self.bodies is not None and getattr(self, "_tile_flags", None) is None:
            self._tile_flags = self.bodies.tile_flags(self.spec)  # this slab's own occupied-tile marks
        b = self.bodies.c_struct(self._tile_flags) if self.bodies is not None else None
        lib = self._lib
        lib.check(lib.load().jaxib_slab_step(
            ops._stream(), self.plan_rows._h, self.plan_cols._h, C.byref(self.c_grid), C.byref(self.c_slab),
            C.byref(b) if b is not None else None, ops._ptr(table), int(nsteps), int(phase_lo), int(phase_hi),
            1 if barriers else 0, C.byref(self.epoch), ops._ptr(self.us), ops._ptr(self.vs)), "slab_step")

    def barrier_timed_out(self) -> bool:
        """True if a barrier of this slab ever gave up waiting for a peer (the results are then invalid)."""
        return bool(self.flags[64].item() != 0)

    def check(self):
        """Raises if a flag barrier gave up: the fields of this and every later step are then built from stale
        or partial peer data.  Called wherever fields are handed back to the host (synchronises)."""
        if self.barrier_timed_out():
            raise RuntimeError(
                f"slab {self.lay.rank}: a peer-memory flag barrier timed out at epoch {int(self.flags[64].item())} "
                "(a peer rank was more than JAXIB_SLAB_TIMEOUT_MS late or died); the fields are invalid")

    def clear_error(self):
        self.flags[64] = 0


def _peer_ranks(spec, particles_bodies, world, halo, device, virtual, group):
    """Allocates the symmetric blocks and returns (ranks, handle)."""
    bodies = particles_bodies
    M = bodies.n_markers if bodies is not None else 0
    if virtual:
        lays = [SlabLayout(spec.nx, spec.ny, world, r, halo) for r in range(world)]
        block = PeerBlock(lays[0], M)
        locals_ = [torch.zeros(block.nbytes, dtype=torch.uint8, device=device) for _ in range(world)]
        ptrs = [t.data_ptr() for t in locals_]
        return [PeerRank(spec, lays[r], bodies, device, block, locals_[r], ptrs) for r in range(world)], None
    import torch.distributed._symmetric_memory as symm_mem
    grp = group if group is not None else dist.group.WORLD
    rank = dist.get_rank(grp)
    lay = SlabLayout(spec.nx, spec.ny, world, rank, halo)
    block = PeerBlock(lay, M)
    local = symm_mem.empty(block.nbytes, dtype=torch.uint8, device=device)
    hdl = symm_mem.rendezvous(local, grp)
    local.zero_()
    torch.cuda.synchronize(device)
    dist.barrier(group=grp)  # every block is zeroed (flags!) before anyone stores into a peer
    return [PeerRank(spec, lay, bodies, device, block, local, [int(p) for p in hdl.buffer_ptrs])], hdl


class SlabStepper:
    """One simulation on `world` slabs.  `virtual=True` keeps all slabs in this process (one GPU, tests
    and single-GPU comparison); otherwise this process is one rank of the torch.distributed group."""

    def __init__(self, spec: ops.GridSpec, particles=None, world: Optional[int] = None, halo: Optional[int] = None,
                 device="cuda", virtual: bool = False, group=None, peer: bool = False):
        from . import engine  # local import: engine imports ops as well
        if not torch.cuda.is_available():
            raise RuntimeError("jax_ib_b200 needs a CUDA device: there is no CPU fallback")
        self.spec = spec
        self.device = torch.device(device)
        self.particles = particles
        self.halo = 2 * spec.ib_window + 4 if halo is None else halo
        if self.halo < 2 * spec.ib_window + 4:
            raise ValueError(f"halo {self.halo} < 2 * ib_window + 4: marker windows would leave the slab")
        bodies = engine.get_bodies(particles, self.device) if particles is not None else None
        self.peer = bool(peer)
        self.virtual = bool(virtual) or not dist.is_initialized()
        if self.virtual:
            self.world = int(world or 1)
            self.comm = LocalComm(self.world)
            my = range(self.world)
        else:
            self.comm = DistComm(group)
            self.world = self.comm.world
            my = [self.comm.rank]
        if self.peer:
            # peer-memory mode (csrc/slab.cu): the exchanges are NVLink stores inside the kernels
            self.ranks, self._symm = _peer_ranks(spec, bodies, self.world, self.halo, self.device, self.virtual, group)
        else:
            self.ranks = [SlabRank(spec, SlabLayout(spec.nx, spec.ny, self.world, r, self.halo), bodies, self.device)
                          for r in my]
        self.time = np.float32(0.0)

    def load(self, u, v, p, t0=0.0):
        for r in self.ranks:
            r.load_global(u, v, p)
            if self.peer:
                r.clear_error()
        self.time = np.float32(t0)
        self._synced = False

    def _host_barrier(self):
        """Before the first peer-memory block after a load: every rank has finished its host-side set-up
        (first torch.func trace of the kinematic laws, uploads), so no rank spins on a peer that is still
        seconds away from launching."""
        if self.peer and not self.virtual and not getattr(self, "_synced", False):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.comm.group)
            self._synced = True

    def check(self):
        if self.peer:
            for r in self.ranks:
                r.check()

    def advance(self, nsteps: int, profile: bool = False):
        """`nsteps` steps in place; returns the body centres after the block (host, float32) or None.
        profile=True brackets every phase with CUDA events (summed ms in self.last_profile; synchronises)."""
        table, centers = self.schedule(nsteps)
        self.run(table, nsteps, profile)
        return None if centers is None else centers[-1]

    def schedule(self, nsteps: int):
        """Host part of a block (as engine.FusedStepper._schedule): evaluates the kinematic laws for the
        next `nsteps` steps, uploads the [nsteps, 1, B, 8] table and advances the clock / body centres."""
        table = centers = None
        if self.particles is not None:
            tab, times, centers = kinematics.body_schedule(self.particles, float(self.time), self.spec.dt, nsteps)
            table = torch.from_numpy(np.ascontiguousarray(tab[:, None])).to(self.device)  # [nsteps, 1, B, 8]
            self._seam_idle = self._no_marker_near_the_seam(tab)
            self._body_ranges = {r.lay.rank: self._bodies_near(r.lay, tab) for r in self.ranks} if self.peer else {}
            self._hints_for = table.data_ptr()  # the hints describe THIS table: run() ignores them for any other
            self.time = times[-1]
            self.particles = dataclasses.replace(self.particles, particle_center=centers[-1])
        else:
            for _ in range(nsteps):
                self.time = np.float32(self.time + np.float32(self.spec.dt))
        return table, centers

    def _no_marker_near_the_seam(self, tab) -> bool:
        bodies = self.ranks[0].bodies
        if bodies is None:
            return True
        return seam_is_idle(np.asarray(tab)[:, :, 2], bodies.max_radius, self.spec.ib_window, self.spec.step[0],
                            self.spec.lower[0], self.spec.nx)

    def _bodies_near(self, lay, tab):
        bodies = self.ranks[0].bodies
        if bodies is None:
            return None
        return bodies_near(np.asarray(tab)[:, :, 2], bodies.max_radius, self.spec.ib_window, self.spec.step[0],
                           self.spec.lower[0], lay.own_lo, lay.own_hi, lay.rank == 0, lay.rank == lay.world - 1)

    def run(self, table, nsteps: int, profile: bool = False):
        """Device part of a block: enqueues `nsteps` steps on the current stream."""
        hinted = table is not None and getattr(self, "_hints_for", None) == table.data_ptr()
        idle = hinted and bool(getattr(self, "_seam_idle", False))
        ranges = getattr(self, "_body_ranges", {}) if hinted else {}
        if self.peer:
            tab = None if table is None else table.view(nsteps, -1, 8)
            if not self.virtual and profile:  # measurement only: one call per phase, bracketed by events
                names = ("predict_force", "spread_rows_fwd_sweepA+barrier", "cols_sweepB", "rows_inv+barrier")
                evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(nsteps)]
                self._host_barrier()
                for n in range(nsteps):
                    evs[n][0].record()
                    for ph in range(4):
                        self.ranks[0].run(None if tab is None else tab[n:n + 1], 1, ph, ph, seam_idle=idle,
                                          body_range=ranges.get(self.ranks[0].lay.rank))
                        evs[n][ph + 1].record()
                torch.cuda.synchronize()
                self.last_profile = {nm: sum(e[k].elapsed_time(e[k + 1]) for e in evs) for k, nm in enumerate(names)}
                self.last_profile["step"] = sum(e[0].elapsed_time(e[4]) for e in evs)
                return
            if not self.virtual:
                self._host_barrier()
                self.ranks[0].run(tab, nsteps, seam_idle=idle, body_range=ranges.get(self.ranks[0].lay.rank))  # one C call: 4 phases x nsteps with flag barriers
            else:  # several slabs of one GPU: lockstep, phase by phase (a flag barrier would deadlock here)
                for n in range(nsteps):
                    for ph in range(4):
                        for r in self.ranks:
                            r.run(None if tab is None else tab[n:n + 1], 1, ph, ph, barriers=False, seam_idle=idle,
                                  body_range=ranges.get(r.lay.rank))
            self.last_profile = None
            return
        prof = None
        if profile:
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(PHASES) + 1)] for _ in range(nsteps)]
        for n in range(nsteps):
            state = None if table is None else table[n]
            if profile:
                evs[n][0].record()
                _step(self.ranks, self.comm, state, mark=lambda k, e=evs[n]: e[k + 1].record())
            else:
                _step(self.ranks, self.comm, state)
        if profile:
            torch.cuda.synchronize()
            prof = {name: sum(e[k].elapsed_time(e[k + 1]) for e in evs) for k, name in enumerate(PHASES)}
            prof["step"] = sum(e[0].elapsed_time(e[-1]) for e in evs)
        self.last_profile = prof

    def owned_fields(self):
        """[(own_lo, u_rows, v_rows, p_rows)] for the slabs held by this process."""
        self.check()
        return [(r.lay.own_lo, r.owned(r.u), r.owned(r.v), r.owned(r.p)) for r in self.ranks]

    def gather(self):
        """Full (u, v, p) on every rank (tests / output; not part of the timed path)."""
        parts = self.owned_fields()
        if isinstance(self.comm, LocalComm):
            return tuple(torch.cat([q[k] for q in parts], dim=0) for k in (1, 2, 3))
        outs = []
        for k in (1, 2, 3):
            mine = parts[0][k].contiguous()
            bufs = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(bufs, mine, group=self.comm.group)
            outs.append(torch.cat(bufs, dim=0))
        return tuple(outs)
