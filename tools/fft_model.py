"""CPU model of the index algebra of the team FFT kernels (csrc/poisson_team.cu): same thread -> element maps,
same padded shared-memory layout, same twiddle tables and digit-reversed positions, executed with NumPy.
Used here (no GPU) to validate the data movement before the kernels run on a B200, and to count the
shared-memory wavefronts of every warp access (bank conflicts).

    python tools/fft_model.py
"""
import itertools

import numpy as np


def phys(p, pad_shift=4):
    return p + (p >> pad_shift)


def wavefronts(addr_words, width_words):
    """Shared-memory wavefronts of one warp access: addr_words[lane] = first 4-byte word, each lane reads
    `width_words` consecutive words.  Hardware model: 32 banks x 4 bytes; a wavefront serves at most one
    distinct 128-byte... simplified: lanes are served in groups (32/width lanes per phase), conflicts =
    max number of distinct words per bank within a phase."""
    lanes = len(addr_words)
    per_phase = max(1, 32 // width_words)
    total = 0
    for ph in range(0, lanes, per_phase):
        banks = {}
        for a in addr_words[ph:ph + per_phase]:
            for w in range(width_words):
                banks.setdefault((a + w) % 32, set()).add(a + w)
        total += max(len(s) for s in banks.values())
    return total


def pos_of_freq(k, radices):
    n = int(np.prod(radices))
    p, ln = 0, n
    for r in radices:
        ln //= r
        p += (k % r) * ln
        k //= r
    return p


class ColModel:
    """K5: column transform of CS column slots, N = R1 R2 R3 rows, team size T = N / 16."""

    def __init__(self, R1, R2, R3, CS=8, pad_shift=4):
        self.R = (R1, R2, R3)
        self.N = R1 * R2 * R3
        self.T = self.N // 16
        self.CS = CS
        self.pad_shift = pad_shift
        self.colpitch = phys(self.N, pad_shift) + 2  # == 2 (mod 16): the G-mapping stores of 8 columns x 2 rows tile the banks
        self.W = np.exp(-2j * np.pi * np.arange(self.N) / self.N)
        self.conflicts = {}

    def ph(self, p):
        return phys(p, self.pad_shift)

    def _log(self, name, addrs_elems):
        """addrs_elems: [threads] element (8-byte) addresses of one access instruction, thread-major."""
        a = np.asarray(addrs_elems)
        w = 0
        nw = 0
        for s in range(0, len(a), 32):
            w += wavefronts(list(2 * a[s:s + 32]), 2)
            nw += 1
        self.conflicts.setdefault(name, []).append(w / nw)

    def run(self, x, lam):
        """x: [N, ncols] complex; lam: [N] eigenvalue of POSITION p (already digit-reversed order) summed
        with the column eigenvalue outside; returns unnormalised inverse(D * forward(x))."""
        R1, R2, R3 = self.R
        N, T, CS = self.N, self.T, self.CS
        ncols = x.shape[1]
        sm = np.zeros(CS * self.colpitch, dtype=np.complex128)
        sub1 = N // R1
        nbA = 16 // R1  # butterflies per thread in pass A
        threads = CS * T
        # ---- pass A forward, G mapping: t -> (c = t % CS, bt = t // CS); butterflies b = bt + T * j
        for j in range(nbA):
            st_addr = [[] for _ in range(R1)]
            for t in range(threads):
                c, bt = t % CS, t // CS
                b = bt + T * j
                v = np.array([x[b + sub1 * r, c] if c < ncols else 0 for r in range(R1)])
                V = np.fft.fft(v)
                for q in range(R1):
                    a = c * self.colpitch + self.ph(b + sub1 * q)
                    sm[a] = V[q] * self.W[(b * q) % N]
                    st_addr[q].append(a)
            for q in range(R1):
                self._log("A.sts", st_addr[q])
        # ---- team passes: t -> (team = t // T, b = t % T)
        sub2 = sub1 // R2
        nbB = 16 // R2
        for j in range(nbB):
            ld = [[] for _ in range(R2)]
            for t in range(threads):
                team, b = t // T, t % T
                bf = b + T * j
                blk, b2 = bf // sub2, bf % sub2
                p0 = blk * sub1 + b2
                idx = [team * self.colpitch + self.ph(p0 + sub2 * r) for r in range(R2)]
                V = np.fft.fft(sm[idx])
                for q in range(R2):
                    sm[idx[q]] = V[q] * self.W[(b2 * q * R1) % N]
                    ld[q].append(idx[q])
            for q in range(R2):
                self._log("B.lds", ld[q])
        nbC = 16 // R3
        for j in range(nbC):
            ld = [[] for _ in range(R3)]
            for t in range(threads):
                team, b = t // T, t % T
                bf = b + T * j
                p0 = bf * R3
                idx = [team * self.colpitch + self.ph(p0 + r) for r in range(R3)]
                V = np.fft.fft(sm[idx])
                V = V * lam[p0:p0 + R3, team] if team < ncols else V
                sm[idx] = np.fft.ifft(V) * R3  # unnormalised inverse
                for q in range(R3):
                    ld[q].append(idx[q])
            for q in range(R3):
                self._log("C.lds", ld[q])
        for j in range(nbB):
            for t in range(threads):
                team, b = t // T, t % T
                bf = b + T * j
                blk, b2 = bf // sub2, bf % sub2
                p0 = blk * sub1 + b2
                idx = [team * self.colpitch + self.ph(p0 + sub2 * r) for r in range(R2)]
                v = sm[idx] * np.conj(self.W[(b2 * np.arange(R2) * R1) % N])
                sm[idx] = np.fft.ifft(v) * R2
        out = np.zeros_like(x)
        for j in range(nbA):
            ld = [[] for _ in range(R1)]
            for t in range(threads):
                c, bt = t % CS, t // CS
                b = bt + T * j
                idx = [c * self.colpitch + self.ph(b + sub1 * r) for r in range(R1)]
                v = sm[idx] * np.conj(self.W[(b * np.arange(R1)) % N])
                V = np.fft.ifft(v) * R1
                for r in range(R1):
                    ld[r].append(idx[r])
                if c < ncols:
                    for r in range(R1):
                        out[b + sub1 * r, c] = V[r]
            for r in range(R1):
                self._log("A'.lds", ld[r])
        return out


def check_cols(R1, R2, R3, CS, ncols, pad_shift=4):
    m = ColModel(R1, R2, R3, CS, pad_shift)
    N = m.N
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, ncols)) + 1j * rng.standard_normal((N, ncols))
    lam_k = -(1.0 + rng.random(N))                    # "eigenvalue" of frequency k
    lam_c = -rng.random(ncols)
    pos = np.array([pos_of_freq(k, (R1, R2, R3)) for k in range(N)])
    lam_pos = np.empty(N)
    lam_pos[pos] = lam_k
    D = 1.0 / (lam_pos[:, None] + lam_c[None, :])    # by position
    got = m.run(x, D)
    want = np.fft.ifft(np.fft.fft(x, axis=0) / (lam_k[:, None] + lam_c[None, :]), axis=0) * N
    err = np.abs(got - want).max() / np.abs(want).max()
    conf = {k: round(float(np.mean(v)), 2) for k, v in m.conflicts.items()}
    print(f"cols N={N} radices=({R1},{R2},{R3}) CS={CS} ncols={ncols} pad>>{pad_shift}: err={err:.2e} "
          f"wavefronts per warp access (ideal 2.0): {conf}")
    return err


if __name__ == "__main__":
    check_cols(8, 8, 8, 8, 7)
    check_cols(16, 8, 8, 8, 8)
    check_cols(16, 16, 8, 8, 7)
    check_cols(16, 16, 16, 4, 4)


# ---------------------------------------------------------------------------------------------- rows
def make_pad(kind):
    if kind == "p4":
        return lambda p: p + (p >> 4)
    if kind == "p3":
        return lambda p: p + (p >> 3)
    if kind == "p46":
        return lambda p: p + (p >> 4) + (p >> 6)
    if kind == "p47":
        return lambda p: p + (p >> 4) + (p >> 7)
    if kind == "p36":
        return lambda p: p + (p >> 3) + (p >> 6)
    if kind == "p48":
        return lambda p: p + (p >> 4) + (p >> 8)
    if kind == "p5":
        return lambda p: p + (p >> 5)
    if kind == "p58":
        return lambda p: p + (p >> 5) + (p >> 8)
    raise ValueError(kind)


class RowModel:
    """K4 / K6: packed-real row transform of length 2N by one team of T = N / 16 threads."""

    def __init__(self, R1, R2, R3, pad="p4"):
        self.R = (R1, R2, R3)
        self.N = R1 * R2 * R3
        self.T = self.N // 16
        self.pad = make_pad(pad)
        self.W = np.exp(-2j * np.pi * np.arange(self.N) / self.N)
        self.W2 = np.exp(-2j * np.pi * np.arange(2 * self.N) / (2 * self.N))
        self.conf = {}

    def _log(self, name, addrs):
        a = np.asarray(addrs)
        w = nw = 0
        for s in range(0, len(a), 32):
            w += wavefronts(list(2 * a[s:s + 32]), 2)
            nw += 1
        self.conf.setdefault(name, []).append(w / nw)

    def pos(self, k):
        return pos_of_freq(k, self.R)

    def forward(self, xrow):
        """xrow: real [2N] -> half spectrum X[0..N] (as numpy rfft)."""
        R1, R2, R3 = self.R
        N, T = self.N, self.T
        z = xrow[0::2] + 1j * xrow[1::2]
        sub1, sub2 = N // R1, N // R1 // R2
        sm = np.zeros(self.pad(N) + 8, dtype=np.complex128)
        for j in range(16 // R1):
            st = [[] for _ in range(R1)]
            for b in range(T):
                bb = b + T * j
                V = np.fft.fft(z[bb + sub1 * np.arange(R1)])
                for q in range(R1):
                    a = self.pad(bb + sub1 * q)
                    sm[a] = V[q] * self.W[(bb * q) % N]
                    st[q].append(a)
            for q in range(R1):
                self._log("A.sts", st[q])
        for j in range(16 // R2):
            ld = [[] for _ in range(R2)]
            for b in range(T):
                bf = b + T * j
                blk, b2 = bf // sub2, bf % sub2
                idx = [self.pad(blk * sub1 + b2 + sub2 * r) for r in range(R2)]
                V = np.fft.fft(sm[idx])
                for q in range(R2):
                    sm[idx[q]] = V[q] * self.W[(b2 * q * R1) % N]
                    ld[q].append(idx[q])
            for q in range(R2):
                self._log("B.lds", ld[q])
        for j in range(16 // R3):
            ld = [[] for _ in range(R3)]
            for b in range(T):
                p0 = (b + T * j) * R3
                idx = [self.pad(p0 + r) for r in range(R3)]
                sm[idx] = np.fft.fft(sm[idx])
                for q in range(R3):
                    ld[q].append(idx[q])
            for q in range(R3):
                self._log("C.lds", ld[q])
        X = np.zeros(N + 1, dtype=np.complex128)
        trips = (N // 2) // T
        for m in range(trips + 1):
            la, lb = [], []
            for b in range(T):
                k = b + T * m
                if k > N // 2:
                    continue
                if k == 0:
                    zz = sm[self.pad(0)]
                    X[0] = zz.real + zz.imag
                    X[N] = zz.real - zz.imag
                    la.append(self.pad(0)); lb.append(self.pad(0))
                else:
                    km = N - k
                    a1, a2 = self.pad(self.pos(k)), self.pad(self.pos(km))
                    zk, zm = sm[a1], np.conj(sm[a2])
                    e = 0.5 * (zk + zm)
                    o = -0.5j * (zk - zm)
                    tt = self.W2[k] * o
                    X[k] = e + tt
                    X[km] = np.conj(e - tt)
                    la.append(a1); lb.append(a2)
            if m < trips:
                self._log("U.lds_k", la)
                self._log("U.lds_m", lb)
        return X

    def inverse(self, X):
        """half spectrum X[0..N] -> real row [2N] (unnormalised: numpy irfft * N)."""
        R1, R2, R3 = self.R
        N, T = self.N, self.T
        sub1, sub2 = N // R1, N // R1 // R2
        sm = np.zeros(self.pad(N) + 8, dtype=np.complex128)
        for m in range((N // 2) // T + 1):
            for b in range(T):
                k = b + T * m
                if k > N // 2:
                    continue
                if k == 0:
                    x0, xn = X[0].real, X[N].real
                    sm[self.pad(0)] = 0.5 * (x0 + xn) + 0.5j * (x0 - xn)
                else:
                    km = N - k
                    xk, xm = X[k], np.conj(X[km])
                    e = 0.5 * (xk + xm)
                    o = 0.5 * (xk - xm) * np.conj(self.W2[k])
                    sm[self.pad(self.pos(k))] = e + 1j * o
                    if km != k:
                        sm[self.pad(self.pos(km))] = np.conj(e) + 1j * np.conj(o)
        for j in range(16 // R3):
            for b in range(T):
                p0 = (b + T * j) * R3
                idx = [self.pad(p0 + r) for r in range(R3)]
                sm[idx] = np.fft.ifft(sm[idx]) * R3
        for j in range(16 // R2):
            for b in range(T):
                bf = b + T * j
                blk, b2 = bf // sub2, bf % sub2
                idx = [self.pad(blk * sub1 + b2 + sub2 * r) for r in range(R2)]
                v = sm[idx] * np.conj(self.W[(b2 * np.arange(R2) * R1) % N])
                sm[idx] = np.fft.ifft(v) * R2
        z = np.zeros(N, dtype=np.complex128)
        for j in range(16 // R1):
            for b in range(T):
                bb = b + T * j
                idx = [self.pad(bb + sub1 * r) for r in range(R1)]
                v = sm[idx] * np.conj(self.W[(bb * np.arange(R1)) % N])
                z[bb + sub1 * np.arange(R1)] = np.fft.ifft(v) * R1
        out = np.empty(2 * N)
        out[0::2], out[1::2] = z.real, z.imag
        return out


def check_rows(R1, R2, R3, pad):
    m = RowModel(R1, R2, R3, pad)
    N = m.N
    x = np.random.default_rng(1).standard_normal(2 * N)
    X = m.forward(x)
    want = np.fft.rfft(x)
    e1 = np.abs(X - want).max() / np.abs(want).max()
    back = m.inverse(want)
    e2 = np.abs(back - x * N).max() / np.abs(x * N).max()
    conf = {k: round(float(np.mean(v)), 2) for k, v in m.conf.items()}
    print(f"rows N={N} ({R1},{R2},{R3}) pad={pad}: fwd err={e1:.1e} inv err={e2:.1e} wavefronts/access: {conf}")


def rows_main():
    for radices in ((8, 8, 8), (16, 8, 8), (8, 16, 8), (16, 16, 8), (16, 16, 16)):
        for pad in ("p4", "p3", "p46", "p47", "p36", "p48", "p5", "p58"):
            try:
                check_rows(*radices, pad)
            except Exception as e:  # noqa: BLE001
                print(radices, pad, "failed", e)


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "rows":
        rows_main()
