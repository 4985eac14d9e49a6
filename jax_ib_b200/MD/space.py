"""jax_md.space.periodic as the notebooks use it (journal_bearing_demo.ipynb cell 8):
`displacement, shift = space.periodic(side, wrapped=False)`; `shift` does not wrap positions."""
from __future__ import annotations

import numpy as np
import torch


def periodic(side, wrapped: bool = False):
    box = np.broadcast_to(np.asarray(side, dtype=np.float32), (2,)).copy()

    def displacement(Ra, Rb, **kwargs):
        d = Ra - Rb
        b = torch.as_tensor(box, device=d.device)
        return d - b * torch.round(d / b)

    def shift(R, dR, **kwargs):
        out = R + dR
        if wrapped:
            b = torch.as_tensor(box, device=out.device)
            out = torch.remainder(out, b)
        return out

    displacement.box = box
    return displacement, shift
