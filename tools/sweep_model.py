"""NumPy float32 restatement of the column sweep (csrc/poisson_sweep.cu): the x-direction of the pressure solve by two
first-order sweeps per y-wavenumber instead of a column FFT.  Host-side model of the kernel's arithmetic -- 16-row
segments, the 32-lane shuffle scan of the carried values, the cyclic closure, the k = 0 column by two prefix sums, and
the slab-level closure of the decomposed passes A / B -- used by tests/test_cpu_host.py against the oracle's FFT solve
and to measure the round-off before any GPU time is spent:

    python tools/sweep_model.py [n]         # rel-L2 against the float64 solution next to the float32 FFT solve
"""
from __future__ import annotations

import sys

import numpy as np
import scipy.fft as sf

F = np.float32
RS = 16


def fma(a, b, c):
    return (np.asarray(a, F) * np.asarray(b, F) + np.asarray(c, F)).astype(F)


def then(f, g):
    """Affine maps y = e + P x as (e, P): first f, then g."""
    return fma(g[1], f[0], g[0]), (g[1] * f[1]).astype(F)


def carry_scan(e, S, Pseg, Plast, rG, rGl, close, cyclic=True, ain0=None, bin_last=None):
    """e[s], S[s] (segment aggregates, float32 arrays over any trailing shape) -> (Ain[s], Bin[s], total forward map,
    total backward map); 32 lanes own ceil(segs / 32) consecutive segments each, as the kernel's warp."""
    segs = e.shape[0]
    m = -(-segs // 32)
    lanes = [(min(l * m, segs), min(l * m + m, segs)) for l in range(32)]
    zero, one = np.zeros_like(e[0]), np.ones_like(e[0])
    P = lambda s: Plast if s == segs - 1 else Pseg
    G = lambda s: rGl if s == segs - 1 else rG
    f = []
    for s0, s1 in lanes:
        g = (zero, one)
        for s in range(s0, s1):
            g = then(g, (e[s], P(s)))
        f.append(g)
    inc = list(f)
    o = 1
    while o < 32:
        inc = [then(inc[l - o], inc[l]) if l >= o else inc[l] for l in range(32)]
        o *= 2
    tot_e = inc[31][0]
    a0 = (tot_e * close).astype(F) if cyclic else ain0
    Ain, T = np.zeros_like(e), np.zeros_like(e)
    back = []
    for l, (s0, s1) in enumerate(lanes):
        cur = a0 if l == 0 else fma(inc[l - 1][1], a0, inc[l - 1][0])
        ge, W = zero, one
        for s in range(s0, s1):
            Ain[s] = cur
            T[s] = fma(cur, G(s), S[s])
            ge = fma(T[s], W, ge)
            W = (W * P(s)).astype(F)
            cur = fma(P(s), cur, e[s])
        back.append((ge, W))
    dec = list(back)
    o = 1
    while o < 32:
        dec = [then(dec[l + o], dec[l]) if l + o < 32 else dec[l] for l in range(32)]
        o *= 2
    tot_t = dec[0][0]
    bl = (tot_t * close).astype(F) if cyclic else bin_last
    Bin = np.zeros_like(e)
    for l, (s0, s1) in enumerate(lanes):
        b = bl if l == 31 else fma(dec[l + 1][1], bl, dec[l + 1][0])
        for s in range(s1 - 1, s0 - 1, -1):
            Bin[s] = b
            b = fma(P(s), b, T[s])
    return Ain, Bin, tot_e, tot_t


def constants(mu, nx_local, nx_total):
    """Per-column constants as cols_sweep_tables computes them (double, rounded to float32)."""
    rho = 1.0 / (1.0 + 0.5 * mu + np.sqrt(mu + 0.25 * mu * mu))
    segs = -(-nx_local // RS)
    last = nx_local - (segs - 1) * RS
    rg = lambda n: rho * (1.0 - rho ** (2 * n)) / (1.0 - rho * rho)
    return dict(rho=rho.astype(F), Pseg=(rho ** RS).astype(F), Plast=(rho ** last).astype(F),
                close=(1.0 / (1.0 - rho ** nx_total)).astype(F), rG=rg(RS).astype(F), rGl=rg(last).astype(F),
                Pblock=(rho ** nx_local).astype(F), rGblock=rg(nx_local).astype(F), rho64=rho)


def local_sweep(X, rho):
    """Step 1 of the kernel: segment-local forward sweep from zero, e and the Horner sum S of every segment."""
    nx = X.shape[0]
    segs = -(-nx // RS)
    a = np.zeros_like(X)
    e = np.zeros((segs,) + X.shape[1:], X.dtype)
    S = np.zeros_like(e)
    for s in range(segs):
        ap = np.zeros_like(X[0])
        r1 = min(RS, nx - s * RS)
        for r in range(r1):
            ap = fma(rho, ap, X[s * RS + r])
            a[s * RS + r] = ap
        h = np.zeros_like(X[0])
        for r in range(r1 - 1, -1, -1):
            h = fma(rho, h, a[s * RS + r])
        e[s], S[s] = ap, h
    return a, e, S


def backward(a, Ain, Bin, rho):
    """Step 3: carried part of the forward sweep + backward sweep (returns b; Q = gain b)."""
    nx = a.shape[0]
    segs = Ain.shape[0]
    out = np.zeros_like(a)
    r64 = rho.astype(np.float64)
    for s in range(segs):
        r1 = min(RS, nx - s * RS)
        b = Bin[s]
        for r in range(r1 - 1, -1, -1):
            pw = (r64 ** (r + 1)).astype(F)  # the kernel builds these from rho, rho^2, rho^4, rho^8 (<= 4 roundings)
            b = fma(rho, b, fma(pw, Ain[s], a[s * RS + r]))
            out[s * RS + r] = b
    return out


def mean_mode(x, dx):
    """Q[i-1] - 2 Q[i] + Q[i+1] = dx^2 (x - mean x), mean Q = 0: two prefix sums in double (the column-0 CTA)."""
    r = x.astype(np.float64)
    r = (r - r.mean()) * dx * dx
    s = np.cumsum(r)
    s -= s.mean()
    q = np.concatenate([[0.0], np.cumsum(s)[:-1]])
    return q - q.mean()


def solve(rhs, dx, dy, slabs=1):
    """Periodic Poisson solve of `rhs` [nx, ny] the way K4 + the column sweep + K6 do it (the row transforms by scipy's
    float32 FFT).  slabs > 1: the decomposed passes A / B with the slab-level closure."""
    nx, ny = rhs.shape
    X = sf.rfft(rhs.astype(F), axis=1).astype(np.complex64)
    k = np.arange(ny // 2 + 1)
    mu = (2.0 - 2.0 * np.cos(2 * np.pi * k / ny)) * dx * dx / (dy * dy)
    Q = np.zeros_like(X)
    nloc = nx // slabs
    c = constants(mu[1:], nloc, nx)
    for part in (np.real, np.imag):
        Xp = part(X[:, 1:]).astype(F)
        if slabs == 1:
            a, e, S = local_sweep(Xp, c["rho"])
            Ain, Bin, _, _ = carry_scan(e, S, c["Pseg"], c["Plast"], c["rG"], c["rGl"], c["close"])
            b = backward(a, Ain, Bin, c["rho"])
        else:
            blocks = [local_sweep(Xp[d * nloc:(d + 1) * nloc], c["rho"]) for d in range(slabs)]
            zero = np.zeros_like(Xp[0])
            maps = [carry_scan(e, S, c["Pseg"], c["Plast"], c["rG"], c["rGl"], c["close"], False, zero, zero)[2:]
                    for _, e, S in blocks]  # pass A: every slab's (forward, backward) map for zero carried-in values
            b = np.zeros_like(Xp)
            for me in range(slabs):  # pass B on slab `me`: close the ring, then sweep with the carried values
                acc = zero
                for j in range(slabs):
                    acc = fma(c["Pblock"], acc, maps[(me + j) % slabs][0])
                ain0 = (acc * c["close"]).astype(F)
                ain_d, T = ain0, []
                for j in range(slabs):
                    E, T0 = maps[(me + j) % slabs]
                    T.append(fma(c["rGblock"], ain_d, T0))
                    ain_d = fma(c["Pblock"], ain_d, E)
                bacc = zero
                for j in range(slabs):
                    bacc = fma(c["Pblock"], bacc, T[0 if j == 0 else slabs - j])
                binl = (bacc * c["close"]).astype(F)
                a, e, S = blocks[me]
                Ain, Bin, _, _ = carry_scan(e, S, c["Pseg"], c["Plast"], c["rG"], c["rGl"], c["close"], False, ain0, binl)
                b[me * nloc:(me + 1) * nloc] = backward(a, Ain, Bin, c["rho"])
        gain = (-dx * dx * c["rho64"]).astype(F)
        if part is np.real:
            Q[:, 1:] += gain * b
        else:
            Q[:, 1:] += 1j * (gain * b)
    Q[:, 0] = mean_mode(X[:, 0].real, dx).astype(F)
    return sf.irfft(Q, n=ny, axis=1).astype(F)


def solve_fft(rhs, dx, dy, dtype):
    nx, ny = rhs.shape
    lx = (2 * np.cos(2 * np.pi * np.arange(nx) / nx) - 2) / dx ** 2
    ly = (2 * np.cos(2 * np.pi * np.arange(ny // 2 + 1) / ny) - 2) / dy ** 2
    lam = lx[:, None] + ly[None, :]
    with np.errstate(divide="ignore"):
        inv = np.where(np.abs(lam) > 10 * np.finfo(dtype).eps, 1.0 / lam, 0.0)
    if dtype == np.float32:
        return sf.irfft2((sf.rfft2(rhs.astype(F)) * inv.astype(F)).astype(np.complex64), s=rhs.shape).astype(F)
    return np.fft.irfft2(np.fft.rfft2(rhs.astype(np.float64)) * inv, s=rhs.shape)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    rng = np.random.default_rng(0)
    x = np.arange(n) * 0.025
    rhs = (np.sin(2 * np.pi * x / (n * 0.025))[:, None] * np.cos(4 * np.pi * x / (n * 0.025))[None, :]
           + 0.3 * rng.standard_normal((n, n))).astype(F)
    rhs -= rhs.mean()
    q64 = solve_fft(rhs, 0.025, 0.025, np.float64)
    rel = lambda q: float(np.linalg.norm(q - q64) / np.linalg.norm(q64))
    print(f"{n}^2: float32 FFT solve {rel(solve_fft(rhs, 0.025, 0.025, np.float32)):.2e}  column sweep {rel(solve(rhs, 0.025, 0.025)):.2e}"
          f"  4 slabs {rel(solve(rhs, 0.025, 0.025, 4)):.2e}")
