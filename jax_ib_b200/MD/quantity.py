"""jax_md.quantity.force as the notebooks use it: turns an energy function into a force function."""
from __future__ import annotations

import torch

from .. import engine, ops
from .interaction_potential import PairEnergy


def force(energy_fn):
    if isinstance(energy_fn, PairEnergy):
        e = energy_fn

        def pair_force(R, **kwargs):
            Rd = engine.to_device(R)
            dev = Rd.device
            return ops.pair_force(Rd, torch.as_tensor(e.species, device=dev), torch.as_tensor(e.h, device=dev),
                                  torch.as_tensor(e.r0, device=dev), (float(e.box[0]), float(e.box[1])), e.D0, e.alpha, e.k)

        return pair_force

    def autograd_force(R, **kwargs):
        Rr = R.detach().clone().requires_grad_(True)
        return -torch.autograd.grad(energy_fn(Rr, **kwargs), Rr)[0]

    return autograd_force
