"""Device-resident fused stepping: the B200 replacement of the reference's `step_fn`
(time_stepping.py:144-255) and of `funcutils.repeated(step_fn, n)`.

One `FusedStepper` owns a jaxib_plan + scratch for a grid and advances (u, v, p) on the GPU for a
block of steps with a single host round: the host evaluates the user's kinematic callables for the
whole block (kinematics.body_schedule -> a [nsteps, B, 8] float32 table, uploaded once), the kernels
run back to back on the current CUDA stream, and the host-side members of All_Variables (clock in
the velocity BC, body centres, Step_count) are advanced with the reference's float32 arithmetic.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional

import numpy as np
import torch

from . import boundaries, grids, kinematics, ops, particle_class


def effective_ib_window(window: int, step) -> int:
    """The window (in cells, same in x and y) that keeps `window` standard deviations of EVERY Gaussian the reference
    sums: interpolation uses widths (dx, dy) (convolution_functions.py:19-21), spreading a radial Gaussian of width dx
    in both directions (IBM_Force.py:67), so on a grid with dy < dx the y-window must grow by dx/dy.  Isotropic grids
    (every reference configuration) are unchanged.  The kernels take windows up to 16."""
    dx, dy = float(step[0]), float(step[1])
    return min(16, int(math.ceil(window * max(1.0, dx / dy) - 1e-9)))


def grid_spec_from(grid: grids.Grid, dt, viscosity, density, bc_kind, convect, v=None, force=(0.0, 0.0), ib_window=6,
                   batch=1) -> ops.GridSpec:
    ib_window = effective_ib_window(ib_window, grid.step)
    u_wall = v_wall = u_wall_x = v_wall_x = (0.0, 0.0)
    if v is not None and bc_kind in ("channel", "far_field"):
        u_wall = tuple(float(x) for x in v[0].bc.bc_values[1])
        v_wall = tuple(float(x) for x in v[1].bc.bc_values[1])
    if v is not None and bc_kind == "far_field":
        u_wall_x = tuple(float(x) for x in v[0].bc.bc_values[0])
        v_wall_x = tuple(float(x) for x in v[1].bc.bc_values[0])
    return ops.GridSpec(u_wall_x=u_wall_x, v_wall_x=v_wall_x,
        nx=grid.shape[0], ny=grid.shape[1], lower=(grid.domain[0][0], grid.domain[1][0]), step=grid.step, dt=dt,
        nu=None if viscosity is None else viscosity / density, density=density, bc=bc_kind, convect=convect,
        u_wall=u_wall, v_wall=v_wall, force=force, ib_window=ib_window, batch=batch)


def cuda_device(device=None) -> torch.device:
    """The CUDA device work runs on: the tensor's own device if it is a CUDA tensor, else the current one."""
    if not torch.cuda.is_available():
        raise RuntimeError("jax_ib_b200 needs a CUDA device: there is no CPU fallback")
    if device is not None and torch.device(device).type == "cuda":
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device())


def to_device(t, device=None) -> torch.Tensor:
    """float32, contiguous, on the GPU (a no-op for tensors that already are)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t, dtype=np.float32))
    return t.to(cuda_device(t.device if device is None else device), dtype=torch.float32).contiguous()


_PLANS = {}
_BODIES = {}


def get_plan(spec: ops.GridSpec, n_markers: int = 0) -> ops.Plan:
    """Plans (FFT tables + scratch) are cached per (grid spec, device)."""
    dev = cuda_device()
    key = (spec.nx, spec.ny, spec.batch, spec.lower, spec.step, spec.bc, spec.u_wall, spec.v_wall, spec.u_wall_x,
           spec.v_wall_x, dev.index)
    plan = _PLANS.get(key)
    if plan is None or plan.n_markers < n_markers:
        plan = ops.Plan(spec, n_markers, dev)
        _PLANS[key] = plan
    return plan


def get_bodies(particles, device) -> ops.DeviceBodies:
    dev = cuda_device(device)
    # everything the generated body-frame markers depend on: the shape callable, its parameters and the
    # theta grid (a resolution study keeps shape_fn and geometry but changes Grid1d)
    G = particles.Grid
    key = (particles.shape, tuple(np.asarray(particles.geometry_param, dtype=np.float64).ravel()),
           getattr(G, "shape", None), tuple(getattr(G, "domain", ()) or ()), dev.index)
    b = _BODIES.get(key)
    if b is None:
        x0s, y0s = [], []
        grid_p = particles.generate_grid()
        for gp in particles.geometry_param:
            x0, y0 = particles.shape(gp, grid_p)
            x0s.append(np.asarray(x0, dtype=np.float32))
            y0s.append(np.asarray(y0, dtype=np.float32))
        b = ops.DeviceBodies(x0s, y0s, dev)
        _BODIES[key] = b
    return b


def bc_kind_of(v) -> str:
    if boundaries.has_all_periodic_boundary_conditions(*v):
        return "periodic"
    t = v[0].bc.types
    if t[0][0] == boundaries.BCType.PERIODIC and t[1][0] == boundaries.BCType.DIRICHLET:
        return "channel"
    if all(x == boundaries.BCType.DIRICHLET for ax in t for x in ax):
        return "far_field"
    raise NotImplementedError(f"no CUDA path for velocity BC types {t}")


class FusedStepper:
    def __init__(self, spec: ops.GridSpec, particles: Optional[particle_class.particle] = None, perm=None,
                 incremental_pressure: bool = True, md_step: Optional[Callable] = None, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("jax_ib_b200 needs a CUDA device: there is no CPU fallback")
        self.spec = spec
        self.device = torch.device(device)
        self.particles_template = particles
        self.bodies = get_bodies(particles, self.device) if particles is not None else None
        self.plan = ops.Plan(spec, self.bodies.n_markers if self.bodies else 0, self.device)
        self.perm = perm
        self.incremental_pressure = incremental_pressure
        self.md_step = md_step
        self.launches_per_step = 4 + (2 if self.bodies else 0)
        self.tile_flags = self.bodies.tile_flags(spec) if self.bodies is not None else None

    # -- raw device-level block ---------------------------------------------------------------
    def run_block(self, u, v, p, nsteps: int, table: Optional[torch.Tensor], walls=None, table_host=None):
        """In place on CUDA tensors (u, v, p).  walls: host [nsteps, 4] wall values per step or None;
        table_host: the kinematic table as a NumPy array (enables the overlapped IB chain)."""
        self.plan.step(u, v, p, nsteps, self.bodies, table, self.perm, self.incremental_pressure, walls, table_host,
                       self.tile_flags)

    # -- All_Variables-level --------------------------------------------------------------------
    def advance(self, all_variables, nsteps: int, move_bodies: bool = True):
        vel = all_variables.velocity
        dev = self.device
        in_dev = vel[0].data.device
        if self.bodies is not None and all_variables.particles is not None:
            b = get_bodies(all_variables.particles, dev)  # a different particle set than at build time
            if b is not self.bodies:
                self.bodies = b
                self.tile_flags = b.tile_flags(self.spec)
                if b.n_markers > self.plan.n_markers:
                    self.plan = ops.Plan(self.spec, b.n_markers, dev)
        # first touch: copy (functional semantics: the caller's arrays are never modified)
        # (non_blocking: from pinned host arrays the three uploads overlap the host-side kinematics below)
        u = vel[0].data.to(dev, dtype=torch.float32, copy=True, non_blocking=True).contiguous()
        v = vel[1].data.to(dev, dtype=torch.float32, copy=True, non_blocking=True).contiguous()
        p = all_variables.pressure.data.to(dev, dtype=torch.float32, copy=True, non_blocking=True).contiguous()
        t0 = vel[0].bc.time_stamp
        dt = self.spec.dt
        particles = all_variables.particles
        state = all_variables
        table_dev, times, centers = self._schedule(particles, t0, dt, nsteps, move_bodies)
        table_host = self._table_host
        walls = self._wall_table(vel, times, nsteps)
        if self.md_step is None:
            self.run_block(u, v, p, nsteps, table_dev, walls, table_host)
        else:
            md = all_variables.MD_var
            for n in range(nsteps):
                self.run_block(u, v, p, 1, None if table_dev is None else table_dev[n:n + 1],
                               None if walls is None else walls[n:n + 1],
                               None if table_host is None else table_host[n:n + 1])
                md = self.md_step(self._view(all_variables, u, v, p, times[n + 1], particles, md))
            state = dataclasses.replace(state, MD_var=md)
        new_particles = particles
        if particles is not None and move_bodies:
            new_particles = dataclasses.replace(particles, particle_center=_like(particles.particle_center, centers[-1]))
        if in_dev.type == "cpu":  # host buffers in, host buffers out: pinned staging + one sync
            outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (u, v, p)]
            for o, t in zip(outs, (u, v, p)):
                o.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            u, v, p = outs
        out = self._view(state, u, v, p, times[-1], new_particles, state.MD_var)
        # particle_motion.py:59: the step counter lives in the position update
        return dataclasses.replace(out, Step_count=all_variables.Step_count + (nsteps if move_bodies else 0))

    # -- rollout: funcutils.trajectory(funcutils.repeated(step_fn, inner), outer) on the device -----------------
    def rollout(self, all_variables, inner_steps: int, outer_steps: int, start_with_input: bool = False,
                move_bodies: bool = True, use_graph: bool = True):
        """The reference's driver loop (Flapping_Demo.ipynb:225-247: `trajectory(repeated(step_fn, inner), outer)`,
        two nested lax.scan's) with ONE host round: the kinematic table of all inner * outer steps is evaluated and
        uploaded once, the inner block is captured in a CUDA graph and replayed per outer step (the launch-bound
        regime of the 300^2 / 600^2 demo grids), and the snapshots go into a device-resident ring
        [outer, 3, nx, ny] that is copied to the host once at the end when the input lived on the host.
        Returns (final All_Variables, [snapshot All_Variables] * outer) like funcutils.trajectory."""
        if self.md_step is not None:
            raise NotImplementedError("rollout with a tracer sub-step: use funcutils.trajectory (per-step host hook)")
        vel = all_variables.velocity
        dev = self.device
        in_dev = vel[0].data.device
        u = vel[0].data.to(dev, dtype=torch.float32, copy=True, non_blocking=True).contiguous()
        v = vel[1].data.to(dev, dtype=torch.float32, copy=True, non_blocking=True).contiguous()
        p = all_variables.pressure.data.to(dev, dtype=torch.float32, copy=True, non_blocking=True).contiguous()
        n_total = inner_steps * outer_steps
        particles = all_variables.particles
        table_dev, times, centers = self._schedule(particles, vel[0].bc.time_stamp, self.spec.dt, n_total, move_bodies)
        walls = self._wall_table(vel, times, n_total)
        ring = torch.empty((outer_steps, 3) + tuple(u.shape), dtype=torch.float32, device=dev)
        graph = block = None
        if use_graph and walls is None and outer_steps > 1:
            try:  # one un-captured block first: every lazy one-time initialisation of the library happens here
                block = None if table_dev is None else table_dev[:inner_steps].clone()
                wu, wv, wp = u.clone(), v.clone(), p.clone()
                self.run_block(wu, wv, wp, inner_steps, block)
                torch.cuda.current_stream().synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.run_block(u, v, p, inner_steps, block)
            except Exception:  # capture not possible on this driver / stream state: plain launches
                graph = None
                torch.cuda.synchronize()
                u.copy_(vel[0].data); v.copy_(vel[1].data); p.copy_(all_variables.pressure.data)
        self.last_rollout_used_graph = graph is not None
        if graph is not None:  # the capture itself did not execute: u, v, p still hold the input
            pass
        for o in range(outer_steps):
            a, b = o * inner_steps, (o + 1) * inner_steps
            if start_with_input:
                ring[o, 0].copy_(u); ring[o, 1].copy_(v); ring[o, 2].copy_(p)
            if graph is not None:
                if block is not None:
                    block.copy_(table_dev[a:b])
                graph.replay()
            else:
                self.run_block(u, v, p, inner_steps, None if table_dev is None else table_dev[a:b],
                               None if walls is None else walls[a:b])
            if not start_with_input:
                ring[o, 0].copy_(u); ring[o, 1].copy_(v); ring[o, 2].copy_(p)
        if in_dev.type == "cpu":  # host in, host out: one pinned staging copy of the ring and of the final state
            hring = torch.empty(ring.shape, dtype=ring.dtype, pin_memory=True)
            hring.copy_(ring, non_blocking=True)
            outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (u, v, p)]
            for o_, t in zip(outs, (u, v, p)):
                o_.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            ring, (u, v, p) = hring, outs

        def state_at(k, fields):  # All_Variables after k steps
            parts = particles
            if particles is not None and move_bodies:
                parts = dataclasses.replace(particles, particle_center=_like(particles.particle_center, centers[k]))
            out = self._view(all_variables, fields[0], fields[1], fields[2], times[k], parts, all_variables.MD_var)
            return dataclasses.replace(out, Step_count=all_variables.Step_count + (k if move_bodies else 0))

        snaps = [state_at((o if start_with_input else o + 1) * inner_steps, ring[o]) for o in range(outer_steps)]
        return state_at(n_total, (u, v, p)), snaps

    def _wall_table(self, vel, times, nsteps):
        """boundaries.py:519-542 update_BC for a block of steps: step 0 sees the wall values stored in the
        incoming BC objects, step n >= 1 the Moving_wall functions (:636-663) evaluated at its own time
        stamp t_n -- exactly what the reference's step n reads after n update_BC calls."""
        if self.spec.bc not in ("channel", "far_field"):
            return None
        far = self.spec.bc == "far_field"
        tab = np.zeros((nsteps, 8), dtype=np.float64)
        tab[:, :4] = [float(x) for x in vel[0].bc.bc_values[1]] + [float(x) for x in vel[1].bc.bc_values[1]]
        if far:
            tab[:, 4:] = [float(x) for x in vel[0].bc.bc_values[0]] + [float(x) for x in vel[1].bc.bc_values[0]]
        fx, fy = vel[0].bc.boundary_fn, vel[1].bc.boundary_fn
        try:
            fns = (fx[2], fx[3], fy[2], fy[3]) + ((fx[0], fx[1], fy[0], fy[1]) if far else ())
        except (TypeError, IndexError):
            return tab  # no per-wall functions (constant-BC factory): the stored values hold for every step
        for n in range(1, nsteps):
            tab[n, :len(fns)] = [float(f(times[n])) for f in fns]
        return tab

    def _schedule(self, particles, t0, dt, nsteps, move_bodies=True):
        f32 = np.float32
        self._table_host = None
        if particles is None:
            times = np.add.accumulate(np.concatenate([[f32(t0)], np.full(nsteps, f32(dt), f32)]).astype(f32), dtype=f32)
            return None, times, None
        table, times, centers = kinematics.body_schedule(particles, t0, dt, nsteps)
        if not move_bodies:  # Updating_Position=None: the reference never touches particle_center (equations.py:143-158)
            table[:, :, 2:4] = centers[0]
        table = np.broadcast_to(table[:, None], (nsteps, self.spec.batch) + table.shape[1:]).copy()
        self._table_host = table
        return torch.from_numpy(table).to(self.device, non_blocking=True), times, centers

    def _view(self, all_variables, u, v, p, t, particles, md):
        vel = all_variables.velocity
        new_v = []
        for comp, data in zip(vel, (u, v)):
            bc = comp.bc
            fns = bc.boundary_fn
            values = bc.bc_values
            try:
                values = ((fns[0](t), fns[1](t)), (fns[2](t), fns[3](t)))
            except TypeError:
                pass
            nbc = boundaries.ConstantBoundaryConditions(values=values, time_stamp=t, types=bc.types, boundary_fn=fns)
            new_v.append(grids.GridVariable(grids.GridArray(data, comp.offset, comp.grid), nbc))
        pr = all_variables.pressure
        new_p = grids.GridVariable(grids.GridArray(p, pr.offset, pr.grid), pr.bc)
        return dataclasses.replace(all_variables, particles=particles, velocity=tuple(new_v), pressure=new_p, MD_var=md)


def _like(template, arr: np.ndarray):
    if isinstance(template, torch.Tensor):
        return torch.from_numpy(np.ascontiguousarray(arr)).to(template.dtype)
    return np.asarray(arr, dtype=np.float32)
