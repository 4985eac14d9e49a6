"""jax_ib/MD/interaction_potential.py:6-22: harmonic wall / Morse tail pair potential with per-species
(h, r0) matrices.  `harmonic_morse_pair(...)` returns the pair ENERGY function as in the reference;
`MD.quantity.force` of it is evaluated by the CUDA all-pairs kernel (csrc/tracer.cu)."""
from __future__ import annotations

import numpy as np
import torch


def harmonic_morse(dr, h=0.5, D0=5.0, alpha=5.0, r0=1.0, k=300.0, **kwargs):
    """interaction_potential.py:6-11."""
    return torch.where(dr < r0, h * k * (dr - r0) ** 2 - D0,
                       D0 * (torch.exp(-2.0 * alpha * (dr - r0)) - 2.0 * torch.exp(-alpha * (dr - r0))))


class PairEnergy:
    def __init__(self, displacement, species, h, D0, alpha, r0, k):
        self.box = getattr(displacement, "box", None)
        if self.box is None:
            raise NotImplementedError("harmonic_morse_pair needs the displacement of MD.space.periodic")
        self.species = np.asarray(species, dtype=np.int32)
        f = lambda x: np.atleast_2d(np.asarray(x, dtype=np.float32))
        ns = int(self.species.max()) + 1
        self.h, self.r0 = np.broadcast_to(f(h), (ns, ns)).copy(), np.broadcast_to(f(r0), (ns, ns)).copy()
        self.D0, self.alpha, self.k = float(D0), float(alpha), float(k)
        self.displacement = displacement

    def __call__(self, R, **kwargs):
        d = self.displacement(R[:, None, :], R[None, :, :])
        r = torch.sqrt((d ** 2).sum(-1) + torch.eye(R.shape[0], device=R.device))
        sp = torch.as_tensor(self.species, device=R.device, dtype=torch.long)
        h = torch.as_tensor(self.h, device=R.device)[sp[:, None], sp[None, :]]
        r0 = torch.as_tensor(self.r0, device=R.device)[sp[:, None], sp[None, :]]
        U = harmonic_morse(r, h, self.D0, self.alpha, r0, self.k)
        return torch.triu(U, 1).sum()


def harmonic_morse_pair(displacement_or_metric, species=None, h=0.5, D0=5.0, alpha=10.0, r0=1.0, k=50.0):
    """interaction_potential.py:15-22."""
    return PairEnergy(displacement_or_metric, species, h, D0, alpha, r0, k)
