"""Named synthetic workloads (BASELINE.json `configs`), as plain data.

Each config is a dict of Python scalars / NumPy arrays so that both the product (engine.py,
bench.py) and the test-side oracle adapter (tests/_oracle_adapter.py) can build their own state
from the same description.  Shapes follow the notebook cells they cite.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def ellipse_markers(geometry_param, ntheta=150):
    """Flapping_Demo.ipynb cell 5 `foil_XY_ELLIPSE`: 2*ntheta - 2 body-frame markers."""
    A, B = F(geometry_param[0]), F(geometry_param[1])
    xt = np.linspace(-A, A, ntheta, dtype=F)
    yt = B / A * np.sqrt(A ** 2 - xt ** 2)
    xt2 = np.linspace(A, -A, ntheta, dtype=F)[1:-1]
    yt2 = -B / A * np.sqrt(A ** 2 - xt2 ** 2)
    return np.append(xt, xt2), np.append(yt, yt2)


def circle_markers(geometry_param, ntheta=80):
    """journal_bearing_demo.ipynb cell 6 `foil_XY_circle`."""
    A = F(geometry_param[0])
    theta = np.linspace(0, F(1.98 * np.pi), ntheta, dtype=F)
    return A * np.cos(theta), A * np.sin(theta)


SHAPES = {"ellipse": ellipse_markers, "circle": circle_markers}


def flap600():
    """G1: Flapping_Demo.ipynb cells 3/5/8 -- 600^2, one flapping ellipse (298 markers), periodic,
    upwind, rho = 1, nu = 0.05, dt = 5e-4."""
    return dict(
        name="flap600", shape=(600, 600), domain=((0.0, 15.0), (0.0, 15.0)), dt=5e-4, density=1.0, viscosity=0.05,
        bc="periodic", convect="upwind", force=(0.0, 0.0), wall_u=(0.0, 0.0), wall_v=(0.0, 0.0),
        bodies=dict(shape="ellipse", geometry=[[0.5, 0.06]], centers=[[15.0 * 0.75, 15.0 / 2]],
                    displacement="harmonic", displacement_param=[[2.8, 0.25]],
                    rotation="harmonic", rotation_param=[[math.pi / 2, math.pi / 4, 0.25, 0.0]]),
        init="zero",
    )


def bearing300():
    """G2: journal_bearing_demo.ipynb cells 4/6 -- 300^2, two concentric rotating circles (80 markers
    each), Fourier kinematics, dt = 8e-5."""
    theta_out = math.pi / 2 * 1.9
    one = np.array([1.0], dtype=F)
    zero = np.array([0.0], dtype=F)
    return dict(
        name="bearing300", shape=(300, 300), domain=((0.0, 5.0), (0.0, 5.0)), dt=8e-5, density=1.0, viscosity=0.05,
        bc="periodic", convect="upwind", force=(0.0, 0.0), wall_u=(0.0, 0.0), wall_v=(0.0, 0.0),
        bodies=dict(shape="circle", geometry=[[0.5, 0.5], [1.0, 1.0]], centers=[[2.5, 2.5], [2.5, 2.5]],
                    displacement="fourier",
                    displacement_param=[[0.0, 0.0, 0.0, one, one, 0.0], [0.0, 0.0, 0.0, one, one, 0.0]],
                    rotation="fourier_normalized",
                    rotation_param=[[math.pi / 2, 1.0, 0.0, zero, zero, theta_out * 3.0, 1.0],
                                    [math.pi / 2, 1.0, 0.0, -zero, zero, theta_out, 1.0]]),
        init="zero",
    )


def _divfree_perturbation(shape, domain, seed, nmodes=16, amp=1e-2):
    """Seeded divergence-free field from a stream function sampled on cell corners:
    u = d(psi)/dy on u-faces, v = -d(psi)/dx on v-faces (discretely divergence free)."""
    rng = np.random.default_rng(seed)
    nx, ny = shape
    (x0, x1), (y0, y1) = domain
    dx, dy = (x1 - x0) / nx, (y1 - y0) / ny
    xc = x0 + np.arange(1, nx + 1) * dx  # corner (i+1, j+1)
    yc = y0 + np.arange(1, ny + 1) * dy
    X, Y = np.meshgrid(xc, yc, indexing="ij")
    psi = np.zeros((nx, ny))
    for _ in range(nmodes):
        kx, ky = rng.integers(1, 5, size=2)
        ph = rng.uniform(0, 2 * np.pi, size=2)
        a = rng.standard_normal()
        psi += a * np.sin(2 * np.pi * kx * (X - x0) / (x1 - x0) + ph[0]) * np.sin(2 * np.pi * ky * (Y - y0) / (y1 - y0) + ph[1])
    u = (psi - np.roll(psi, 1, axis=1)) / dy
    v = -(psi - np.roll(psi, 1, axis=0)) / dx
    s = amp / max(np.abs(u).max(), np.abs(v).max())
    return (u * s).astype(F), (v * s).astype(F)


def channel(shape=(1024, 256), domain=((0.0, 4.0), (0.0, 1.0)), seed=3, force_x=2e-3):
    """G3: taylor_dispersion_demo.ipynb cells 9-11 -- periodic-x / no-slip-y channel, van Leer,
    nu = 1e-3, body force 2e-3 in x, dt = min(0.5 dx / 1, dx^2 / (4 nu)).  The notebook's dt assumes |u| <= 1: a taller
    channel needs a proportionally smaller force (Poiseuille peak = force H^2 / (8 nu)) to stay inside that CFL bound."""
    nx, ny = shape
    dx = (domain[0][1] - domain[0][0]) / nx
    dy = (domain[1][1] - domain[1][0]) / ny
    h = min(dx, dy)
    nu = 1e-3
    dt = min(0.5 * h / 1.0, h ** 2 / (nu * 4))
    return dict(
        name=f"channel{nx}x{ny}", shape=shape, domain=domain, dt=dt, density=1.0, viscosity=nu,
        bc="channel", convect="van_leer", force=(force_x, 0.0), wall_u=(0.0, 0.0), wall_v=(0.0, 0.0),
        bodies=None, init=("poiseuille+noise", seed),
    )


def ib_periodic(n=2048, nbodies_axis=2, seed=4, dx=0.025, flow_amp=0.5):
    """Bench workload (BASELINE.json metric: full IB step at 2048^2 on one B200) and the per-GPU
    tile of G4 `multi8192`: n^2 periodic grid at the Flapping demo's resolution (dx = 0.025),
    nbodies_axis^2 flapping ellipses (298 markers each) on a lattice with phases 2 pi b / B, seeded
    divergence-free initial flow, upwind advection, FFT Poisson."""
    L = n * dx
    B = nbodies_axis * nbodies_axis
    pitch = L / nbodies_axis
    centers = [[(ix + 0.5) * pitch, (iy + 0.5) * pitch] for ix in range(nbodies_axis) for iy in range(nbodies_axis)]
    return dict(
        name=f"ib{n}", shape=(n, n), domain=((0.0, L), (0.0, L)), dt=5e-4, density=1.0, viscosity=0.05,
        bc="periodic", convect="upwind", force=(0.0, 0.0), wall_u=(0.0, 0.0), wall_v=(0.0, 0.0),
        bodies=dict(shape="ellipse", geometry=[[0.5, 0.06]] * B, centers=centers,
                    displacement="harmonic_multi", displacement_param=[[2.8, 0.25]] * B,
                    rotation="harmonic_multi",
                    rotation_param=[[math.pi / 2, math.pi / 4, 0.25, 2 * math.pi * b / B] for b in range(B)]),
        init=("divfree", seed, flow_amp),
    )


def small_periodic(n=64, seed=7, bodies=True, convect="upwind", dx=0.05):
    """Small parity case: n^2 periodic grid, one small ellipse, non-trivial initial flow."""
    L = n * dx
    cfg = dict(
        name=f"small{n}", shape=(n, n), domain=((0.0, L), (0.0, L)), dt=2e-3, density=1.0, viscosity=0.05,
        bc="periodic", convect=convect, force=(0.0, 0.0), wall_u=(0.0, 0.0), wall_v=(0.0, 0.0),
        bodies=dict(shape="ellipse", geometry=[[0.3 * L / 3.2, 0.05 * L / 3.2]], centers=[[L * 0.6, L * 0.5]],
                    displacement="harmonic", displacement_param=[[0.4, 0.5]],
                    rotation="harmonic", rotation_param=[[math.pi / 2, math.pi / 4, 0.5, 0.3]], ntheta=20)
        if bodies else None,
        init=("divfree", seed, 0.5),
    )
    return cfg


def initial_fields(cfg):
    """(u, v, p) float32 arrays for a config."""
    nx, ny = cfg["shape"]
    init = cfg["init"]
    u = np.zeros((nx, ny), F)
    v = np.zeros((nx, ny), F)
    p = np.zeros((nx, ny), F)
    if init == "zero":
        return u, v, p
    kind = init[0]
    if kind == "divfree":
        u, v = _divfree_perturbation(cfg["shape"], cfg["domain"], init[1], amp=init[2])
    elif kind == "poiseuille+noise":
        (y0, y1) = cfg["domain"][1]
        dy = (y1 - y0) / ny
        y = y0 + (np.arange(ny) + 0.5) * dy
        G, nu = cfg["force"][0], cfg["viscosity"]
        H = y1 - y0
        prof = G / (2 * nu) * (y - y0) * (H - (y - y0))
        du, dv = _divfree_perturbation(cfg["shape"], cfg["domain"], init[1], amp=1e-2)
        # keep the wall-normal perturbation compatible with the no-slip walls
        env = np.sin(np.pi * (np.arange(1, ny + 1) * dy) / H) ** 2
        u = (prof[None, :] + du).astype(F)
        v = (dv * env[None, :]).astype(F)
        v[:, -1] = 0.0
    else:
        raise ValueError(init)
    return u, v, p


def body_markers(cfg):
    """Body-frame marker lists [(x0, y0)] per body."""
    b = cfg["bodies"]
    if b is None:
        return []
    fn = SHAPES[b["shape"]]
    kw = {"ntheta": b["ntheta"]} if "ntheta" in b else {}
    return [fn(g, **kw) for g in b["geometry"]]
