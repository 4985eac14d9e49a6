#!/usr/bin/env python
"""bench.py -- Mcell-steps/s of the full IB step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload ib2048]

A "step" is one full immersed-boundary time step (explicit predict -> IB interpolate / force /
spread -> FFT pressure projection) of the `ib2048` workload: a 2048^2 periodic grid at the Flapping
demo's resolution with 4 flapping ellipses (1192 markers), upwind advection (configs.ib_periodic).
N > 1 runs one independent simulation per GPU (ensemble sweep, no data-path collective): weak scaling.

`value`     device-resident throughput, per-step CUDA-event brackets, L2 flushed between steps.
`e2e`       same metric through the public All_Variables -> All_Variables API with HOST (pinned)
            arrays: H2D of (u, v, p) and D2H of the result inside every timed step.
`roofline`  dominant kernel, algorithmic bytes / CUDA-event duration / measured HBM peak.
`cpu_baseline` / `--impl reference`: the NumPy oracle (a faithful restatement of the reference's
            dense algorithm; the reference's JAX path cannot run in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mcell-steps/s (full IB step)"
UNIT = "Mcell-steps/s"
# SURVEY.md 8d / DESIGN.md: algorithmic HBM bytes per cell-step of each grid kernel (fp32)
ALGO_BYTES = {"explicit_predict": 20.0, "rows_fwd": 12.0, "cols": 8.0, "rows_inv": 28.0}
STEP_BYTES = 68.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ib2048")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="skip the 8192^2 slab-decomposed measurement")
    return ap.parse_args()


def workload_cfg(name):
    from jax_ib_b200 import configs

    if name == "ib2048":
        return configs.ib_periodic(2048, 2)
    if name == "ib4096":
        return configs.ib_periodic(4096, 4)
    if name == "ib1024":
        return configs.ib_periodic(1024, 1)
    if name in ("flap600", "ens1024"):  # G1 / G5: the Flapping demo grid; ens1024 = 128 of them per GPU
        return configs.flap600()
    if name == "bearing300":  # G2: journal bearing + 1401 tracers
        return configs.bearing300()
    if name == "channel2048":
        # G3: Taylor-dispersion channel at 2048^2 cells on [0,8]^2 (the notebook's dx, dt).  The force is scaled by
        # 1/H^2 so that the Poiseuille peak stays at the notebook's 0.25 (with 2e-3 the 8-unit-high channel would run
        # at |u| = 16, far outside the CFL bound the notebook's dt assumes: the r2o run went non-finite)
        return configs.channel((2048, 2048), ((0.0, 8.0), (0.0, 8.0)), force_x=2e-3 / 64.0)
    if name.startswith("slab"):  # slab8192 = G4 `multi8192`: (n/2048)^2 flapping ellipses on an n^2 grid
        n = int(name[4:])
        return configs.ib_periodic(n, max(1, n // 1024))  # 8192: 8 x 8 lattice = 64 bodies, 19 072 markers
    raise SystemExit(f"unknown workload {name}")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name, rank, world, device):
    """Everything the timed loops need for one named workload: stepper, batched initial fields, the host-side
    kinematic table builder and an optional per-step hook (tracers)."""
    import numpy as np
    import torch

    from jax_ib_b200 import configs, demo, kinematics, ops

    cfg = workload_cfg(name)
    batch = 128 if name == "ens1024" else 1
    members = [cfg]
    if name == "ens1024":
        # copies of G1 with A0 ~ U[2.0, 3.2], f ~ U[0.15, 0.35], beta ~ U[pi/8, 3 pi/8], phi ~ U[0, 2 pi),
        # seed 1234; member index = rank * 128 + k of the 1024-member sweep
        import math
        rng = np.random.default_rng(1234)
        A0, f = rng.uniform(2.0, 3.2, 1024), rng.uniform(0.15, 0.35, 1024)
        beta, phi = rng.uniform(math.pi / 8, 3 * math.pi / 8, 1024), rng.uniform(0, 2 * math.pi, 1024)
        members = []
        for k in range(batch):
            m = (rank * batch + k) % 1024
            b = dict(cfg["bodies"], displacement_param=[[float(A0[m]), float(f[m])]],
                     rotation_param=[[math.pi / 2, float(beta[m]), float(f[m]), float(phi[m])]])
            members.append(dict(cfg, bodies=b))
    elif world > 1 and cfg["bodies"] is not None and name.startswith("ib"):
        b = dict(cfg["bodies"])  # ensemble sweep: every rank gets its own kinematic phase
        b["rotation_param"] = [[p[0], p[1], p[2], p[3] + 0.37 * rank] for p in b["rotation_param"]]
        cfg = dict(cfg, bodies=b)
        members = [cfg]
    stepper = demo.make_stepper(cfg, batch=batch, device=device)
    f0 = configs.initial_fields(cfg)
    if name == "ens1024":  # a little seeded flow so that the fields are not identically zero
        f0 = configs.initial_fields(dict(cfg, init=("divfree", 11, 0.05)))
    fields = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (batch,) + a.shape) if batch > 1 else a)).to(device)
              for a in f0]

    def table(nsteps):
        if cfg["bodies"] is None:
            return None
        tabs = [kinematics.body_schedule(demo.make_particles(m), 0.0, m["dt"], nsteps)[0] for m in members]
        return np.ascontiguousarray(np.stack(tabs, axis=1))  # [nsteps, batch, B, 8]

    post, extra = None, {}
    if name == "bearing300":
        # journal_bearing_demo.ipynb cells 8-10: 1200 mobile tracers + 200 wall + 1 hub, harmonic-Morse walls, kT = 0
        rng = np.random.default_rng(0)
        n_mob, n_wall = 1200, 200
        th, rr = rng.uniform(0, 2 * np.pi, n_mob), rng.uniform(0.6, 0.9, n_mob)
        tw = np.linspace(0, 2 * np.pi, n_wall)
        pos = np.concatenate([np.stack([2.5 + rr * np.cos(th), 2.5 + rr * np.sin(th)], 1),
                              np.stack([2.5 + np.cos(tw), 2.5 + np.sin(tw)], 1), [[2.5, 2.5]]]).astype(np.float32)
        species = torch.from_numpy(np.concatenate([np.zeros(n_mob), np.ones(n_wall), [2]]).astype(np.int32)).to(device)
        mobile = torch.from_numpy(np.concatenate([np.ones(n_mob), np.zeros(n_wall + 1)]).astype(np.uint8)).to(device)
        r0 = torch.tensor([[0.0, 2 * np.pi / n_wall, 0.5], [2 * np.pi / n_wall, 0, 0], [0.5, 0, 0]], dtype=torch.float32, device=device)
        h = torch.tensor([[0.0, 5, 5], [5, 0, 0], [5, 0, 0]], dtype=torch.float32, device=device)
        R = torch.from_numpy(pos).to(device)
        spec = stepper.spec

        def post(u, v):
            F = ops.pair_force(R, species, h, r0, (5.0, 5.0))
            ops.tracer_step(spec, u, v, R, cfg["dt"], pair=F, is_mobile=mobile)

        extra = {"tracers": int(R.shape[0]), "tracer_launches_per_step": 2}
    return dict(cfg=cfg, batch=batch, stepper=stepper, fields=fields, table=table, post=post, extra=extra)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference(workload, steps, threads, warmup=0):
    """The oracle (NumPy float32 port of the reference's algorithm, dense O(M Nx Ny) IB sums as IBM_Force.py:66-87
    evaluates them) on the SAME workload as the GPU arm, on `threads` host threads: the markers of every body are
    sharded over the threads with the grid shared -- the reference's own parallelisation (jax.pmap over marker
    shards, IBM_Force.py:77-87).  Returns (Mcell-steps/s, seconds per step)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _adapter import oracle_state, oracle_step

    cfg = workload_cfg(workload)
    nx, ny = cfg["shape"]
    st = oracle_state(cfg)
    step = oracle_step(cfg, ib_window=None, ib_threads=threads)
    for _ in range(warmup):
        st = step(st)
    t0 = time.perf_counter()
    for _ in range(steps):
        st = step(st)
    wall = time.perf_counter() - t0
    return nx * ny * steps / wall / 1e6, wall / steps


def reference_sample_text(workload, steps, threads, sec_per_step):
    return (f"{steps} full steps of the {workload} workload itself (same grid, bodies, markers, dt, upwind, FFT Poisson) "
            f"with the NumPy float32 oracle, dense O(M*Nx*Ny) IB sums as in IBM_Force.py:66-87, markers sharded over "
            f"{threads} host threads with the grid shared (the reference's pmap layout, IBM_Force.py:77-87): "
            f"{sec_per_step:.1f} s per step")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 64)
    steps = max(1, min(args.steps, 2))
    warm = min(args.warmup, 1)
    val, sec = cpu_reference(args.workload, steps, threads, warm)
    cfg = workload_cfg(args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "grid": list(cfg["shape"]),
                   "bodies": len(cfg["bodies"]["centers"]) if cfg["bodies"] else 0, "convect": cfg["convect"],
                   "note": "restated-reference CPU (NumPy oracle), not JAX: jax is not installable in this image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": reference_sample_text(args.workload, steps, threads, sec)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ slab arm
def measure_slab(name, K, W, rank, world, device, barrier, max_over_ranks):
    """ONE simulation of the `name` grid on `world` GPUs (strong scaling): slab decomposition along x,
    NCCL all-to-all FFT transpose + halo exchange (jax_ib_b200/slab.py).  At world == 1 the single-GPU
    fused step runs the same grid.  Returns a dict for the JSON line."""
    import numpy as np
    import torch

    from jax_ib_b200 import configs, demo, kinematics, slab

    cfg = workload_cfg(name)
    nx, ny = cfg["shape"]
    cells = nx * ny
    fields = configs.initial_fields(cfg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = None
    if world == 1:
        stepper = demo.make_stepper(cfg, device=device)
        u, v, p = (torch.from_numpy(a).to(device) for a in fields)
        tab, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], W + K)
        tab = torch.from_numpy(np.ascontiguousarray(tab[:, None])).to(device)
        stepper.run_block(u, v, p, W, tab[:W])
        barrier()
        e0.record()
        stepper.run_block(u, v, p, K, tab[W:])
        e1.record()
        barrier()
        finite = bool(torch.isfinite(u).all().item())
        how = "single-GPU fused step (jaxib_step)"
    else:
        def timed(st):
            st.load(*(torch.from_numpy(a) for a in fields))
            st.advance(W)
            table, _ = st.schedule(K)  # as at one GPU: the kinematic table of the block is on the device first
            barrier()
            e0.record()
            st.run(table, K)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1))

        # baseline: the same kernels with NCCL collectives between them (2 all-to-all + pack copies, send/recv)
        st = slab.SlabStepper(demo.make_spec(cfg), demo.make_particles(cfg), device=device)
        L = st.ranks[0].lay
        ms_nccl = None
        if not os.environ.get("JAXIB_BENCH_SKIP_NCCL"):
            ms_nccl = timed(st)
            st.advance(K, profile=True)
            phases = {k: round(max_over_ranks(v) / K * 1e3, 1) for k, v in st.last_profile.items()}
        del st
        torch.cuda.empty_cache()
        # product: exchanges fused into the kernels as NVLink stores to peer-mapped memory + flag barriers
        mode, ms_peer, peer_phases = "peer-memory", None, None
        try:
            st = slab.SlabStepper(demo.make_spec(cfg), demo.make_particles(cfg), device=device, peer=True)
            ms_peer = timed(st)
            st.advance(K, profile=True)
            peer_phases = {k: round(max_over_ranks(v) / K * 1e3, 1) for k, v in st.last_profile.items()}
            if st.ranks[0].barrier_timed_out():
                raise RuntimeError("a flag barrier timed out")
            finite = bool(torch.isfinite(st.ranks[0].u).all().item())
        except Exception as e:  # noqa: BLE001  (no peer mapping on this box: report the NCCL path)
            mode, ms_peer = f"nccl (peer-memory mode unavailable: {type(e).__name__}: {e})"[:200], None
            finite = True
        how = (f"{world} slabs of {L.nloc} rows + {L.halo} halo rows; {mode}: the x-direction of the pressure solve is "
               f"two first-order sweeps per column (no transposes): every slab sweeps its own rows and stores its carried-"
               f"value maps (2 floats per column and part) and its halo rows into the peers' memory over NVLink, 2 flag "
               f"barriers per step, marker forces computed locally from the halo" if ms_peer is not None else
               f"{world} slabs of {L.nloc} rows + {L.halo} halo rows; {mode}")
        e_ms = ms_peer if ms_peer is not None else ms_nccl
    ms = max_over_ranks(e0.elapsed_time(e1)) if world == 1 else e_ms
    # ---- correctness of the decomposition, carried by the driver's own SCALE record: `nv` steps from the initial
    #      state; at N > 1 rank 0 also runs the single-GPU fused step itself and compares: rel-L2 per field (the
    #      column sweep associates the carried values per slab, so slabs and one GPU agree to float32 round-off, not
    #      bit for bit; the transposing paths are bit-identical, tests/test_slab.py) -----------------------------
    import hashlib
    nv = 3
    verify = {"steps": nv}
    try:
        if world == 1:
            u, v, p = (torch.from_numpy(a).to(device) for a in fields)
            tabv, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], nv)
            stepper.run_block(u, v, p, nv, torch.from_numpy(np.ascontiguousarray(tabv[:, None])).to(device))
            got = (u, v, p)
        else:
            st.particles = demo.make_particles(cfg)  # the timed blocks moved the bodies: restart from the initial state
            st.load(*(torch.from_numpy(a) for a in fields))
            st.advance(nv)
            got = st.gather()
            st.check()
        torch.cuda.synchronize()
        h = hashlib.sha1()
        for t in got:
            h.update(t.cpu().numpy().tobytes())
        verify["checksum"] = h.hexdigest()[:16]
        if world > 1 and rank == 0:
            ref_stepper = demo.make_stepper(cfg, device=device)
            u, v, p = (torch.from_numpy(a).to(device) for a in fields)
            tabv, _, _ = kinematics.body_schedule(demo.make_particles(cfg), 0.0, cfg["dt"], nv)
            ref_stepper.run_block(u, v, p, nv, torch.from_numpy(np.ascontiguousarray(tabv[:, None])).to(device))
            verify["bitwise_vs_1gpu"] = bool(all(torch.equal(a, b) for a, b in zip(got, (u, v, p))))
            verify["max_abs_diff_vs_1gpu"] = float(max((a - b).abs().max().item() for a, b in zip(got, (u, v, p))))
            rel = [float(((a - b).norm() / b.norm()).item()) for a, b in zip(got, (u, v, p))]
            verify["rel_l2_vs_1gpu"] = {"u": rel[0], "v": rel[1], "p": rel[2]}
            verify["matches_1gpu"] = bool(rel[0] < 5e-6 and rel[1] < 5e-6 and rel[2] < 5e-5)
            del ref_stepper
    except Exception as e:  # noqa: BLE001
        verify["error"] = f"{type(e).__name__}: {e}"[:200]
    out = {"workload": name, "grid": [nx, ny], "bodies": len(cfg["bodies"]["centers"]), "n_gpus": world,
           "scaling": "strong", "steps": K, "ms_per_step": ms / K, "value": cells * K / (ms * 1e-3) / 1e6,
           "unit": UNIT, "finite": finite, "verify": verify, "markers": int(sum(len(x) for x, _ in configs.body_markers(cfg))),
           "how": how,
           "l2": "state (5 fields + spectrum buffers) per GPU exceeds the 126 MB L2; steps run back to back"}
    if world > 1 and peer_phases:
        out["phase_us_max_over_ranks"] = peer_phases
    if phases:
        out["nccl_path"] = {"ms_per_step": ms_nccl / K, "value": cells * K / (ms_nccl * 1e-3) / 1e6,
                            "phase_us_max_over_ranks": phases}
    return out


# ------------------------------------------------------------------------------------------------ own arm
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from jax_ib_b200 import demo, kinematics, ops

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload.startswith("slab"):
        res = measure_slab(args.workload, args.steps, args.warmup, rank, world, device, barrier, max_over_ranks)
        if rank == 0:
            hbm_peak, peak_src = peaks()
            gbs = STEP_BYTES * res["value"] * 1e6 / 1e9
            line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": res,
                    "step_roofline": {"algorithmic_bytes_per_cell_step": STEP_BYTES, "achieved_gbs": gbs,
                                      "frac_of_aggregate_peak": gbs / (hbm_peak * world), "peak_source": peak_src}}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    wl = make_workload(args.workload, rank, world, device)
    cfg, batch, stepper, post = wl["cfg"], wl["batch"], wl["stepper"], wl["post"]
    nx, ny = cfg["shape"]
    cells = batch * nx * ny
    K, W = args.steps, args.warmup
    u0, v0, p0 = wl["fields"]
    table_host = wl["table"](W + 3 * K)  # the host copy also lets the library overlap the IB chain (JAXIB_IB_OVERLAP=1)
    table = None if table_host is None else torch.from_numpy(table_host).to(device)
    tsl = lambda a, b: (None if table is None else table[a:b], None if table_host is None else table_host[a:b])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # 256 MiB > 126 MB L2

    def run(u, v, p, a, b):
        """steps a..b of the schedule; with a per-step hook (tracers) one launch chain per step"""
        if post is None:
            d, h = tsl(a, b)
            stepper.run_block(u, v, p, b - a, d, table_host=h)
        else:
            for n in range(a, b):
                d, h = tsl(n, n + 1)
                stepper.run_block(u, v, p, 1, d, table_host=h)
                post(u, v)

    u, v, p = u0.clone(), v0.clone(), p0.clone()
    run(u, v, p, 0, W)  # warm-up (untimed)
    # ---- timed region 1: K steps, each bracketed by a CUDA-event pair on the launching stream, 256 MiB
    #      L2 flush before every step outside the brackets (the launches queue up behind the flush) ------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for n in range(K):
        flush.fill_(n & 0xff)
        evs[n][0].record()
        run(u, v, p, W + n, W + n + 1)
        evs[n][1].record()
    barrier()
    ms_steps = sum(a.elapsed_time(b) for a, b in evs)
    # ---- stage breakdown: same K steps again with an event after every kernel (jaxib_step_profile); the
    #      extra events serialise the kernels (no programmatic dependent launch), so its sum is larger ----
    prof = ops.step_profile(stepper.plan, u, v, p, K, stepper.bodies, tsl(W + 2 * K, W + 3 * K)[0], stepper.perm,
                            stepper.incremental_pressure, flush=flush, tile_flags=stepper.tile_flags)
    barrier()
    # ---- timed region 2: K steps back to back (L2 warm, as consecutive time steps really run) ---
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run(u, v, p, W + K, W + 2 * K)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_flushed = max_over_ranks(ms_steps)
    ms_warm = max_over_ranks(e0.elapsed_time(e1))
    finite = bool(torch.isfinite(u).all().item())

    # ---- e2e: the same step through the public API with HOST (pinned) arrays: H2D of (u, v, p) and D2H of the
    #      result inside every timed step ---------------------------------------------------------------------
    Ke = max(3, min(K, 20))
    if batch == 1:
        av = demo.make_all_variables(cfg, device="cpu")
        pin = lambda gv: type(gv)(type(gv.array)(gv.data.pin_memory(), gv.offset, gv.grid), gv.bc)
        av = type(av)(av.particles, tuple(pin(x) for x in av.velocity), pin(av.pressure), av.Drag, av.Step_count, av.MD_var)
        for _ in range(2):
            av = stepper.advance(av, 1)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            av = stepper.advance(av, 1)  # H2D (u, v, p) -> step -> D2H (u, v, p)
        torch.cuda.synchronize()
        e2e_api = "FusedStepper.advance(All_Variables on pinned host memory, 1)"
    else:  # ensemble: the batched block call on pinned host arrays (All_Variables holds one simulation)
        hu, hv, hp = (t.cpu().pin_memory() for t in (u, v, p))
        ou, ov, op = (torch.empty_like(t).pin_memory() for t in (hu, hv, hp))
        barrier()
        t0 = time.perf_counter()
        for n in range(Ke):
            du, dv, dp = (t.to(device, non_blocking=True) for t in (hu, hv, hp))
            d, h = tsl(n, n + 1)
            stepper.run_block(du, dv, dp, 1, d, table_host=h)
            for o, t in zip((ou, ov, op), (du, dv, dp)):
                o.copy_(t, non_blocking=True)
            torch.cuda.synchronize()
        e2e_api = "FusedStepper.run_block on pinned host arrays [batch, nx, ny] (H2D, 1 step, D2H)"
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    # second partitioning: one 8192^2 grid slab-decomposed over the same GPUs (strong scaling)
    slab_res = None
    if args.workload == "ib2048" and not args.no_slab:
        del flush
        torch.cuda.empty_cache()
        slab_res = measure_slab("slab8192", max(5, min(K, 30)), 3, rank, world, device, barrier, max_over_ranks)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = peaks()
    value = world * cells * K / (ms_flushed * 1e-3) / 1e6
    stage_ms = {k: prof[k] / K for k in ops.STAGES}
    dom = max(ALGO_BYTES, key=lambda k: stage_ms[k])
    achieved = ALGO_BYTES[dom] * cells / (stage_ms[dom] * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (ib2048 only)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if args.workload == "ib2048" and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_flushed / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "grid": [nx, ny], "batch_per_gpu": batch,
                   "bodies": len(cfg["bodies"]["centers"]) if cfg["bodies"] else 0,
                   "markers": stepper.bodies.n_markers if stepper.bodies else 0, "convect": cfg["convect"],
                   "poisson": "periodic FFT" if cfg["bc"] == "periodic" else "channel: FFT x DCT-II", **wl["extra"],
                   "ib_window_R": 6,
                   "parallelism": f"ensemble x{world} ({batch} simulation(s) per GPU, no collective)",
                   "l2": "flushed (256 MiB fill) before every timed step; one event pair per step, the flush is outside it",
                   "finite": finite},
        "value_back_to_back": world * cells * K / (ms_warm * 1e-3) / 1e6,
        "ms_per_step_back_to_back": ms_warm / K,
        "step_roofline": {"algorithmic_bytes_per_cell_step": STEP_BYTES,
                          "achieved_gbs": STEP_BYTES * cells * K / (ms_flushed * 1e-3) / 1e9,
                          "frac": STEP_BYTES * cells * K / (ms_flushed * 1e-3) / 1e9 / hbm_peak,
                          "frac_back_to_back": STEP_BYTES * cells * K / (ms_warm * 1e-3) / 1e9 / hbm_peak},
        "stage_us": {k: round(v * 1e3, 2) for k, v in stage_ms.items()},
        "stage_note": "per-kernel event brackets (separate pass): they serialise the launches, so sum(stage_us) > ms_per_step",
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALGO_BYTES[dom] * cells},
        "clocks": clocks,
        "gpu_launches": (stepper.launches_per_step + wl["extra"].get("tracer_launches_per_step", 0) +
                         (1 if os.environ.get("JAXIB_IB_OVERLAP", "0") == "1" else 0)) * K,
        "e2e": {"value": world * cells * Ke / e2e_s / 1e6, "unit": UNIT, "steps": Ke,
                "h2d_bytes_per_step": 3 * cells * 4, "d2h_bytes_per_step": 3 * cells * 4, "api": e2e_api},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = min(os.cpu_count() or 1, 64)
        val, sec = cpu_reference(args.workload, 1, threads)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": reference_sample_text(args.workload, 1, threads, sec) +
                                ("; ONE member of the ensemble" if batch > 1 else "") +
                                ("; fluid step only (no tracer sub-step)" if post is not None else "")}
    if slab_res is not None:
        line["slab"] = slab_res
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: spawn it the way the driver would
        import subprocess

        port = os.environ.get("MASTER_PORT", "29541")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
