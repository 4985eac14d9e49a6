"""`repeated` / `trajectory` with the semantics of jax_cfd.base.funcutils (un-vendored upstream the
reference's notebooks call, Flapping_Demo.ipynb:225-247).  A fused step function produced by
equations.semi_implicit_navier_stokes_timeBC exposes `.advance(x, n)` and `.rollout(x, inner, outer)`:
`repeated` then runs a whole inner loop per host round trip, and `trajectory(repeated(step_fn, inner), outer)`
-- the reference's driver -- runs ENTIRELY on the device (engine.FusedStepper.rollout: one kinematic table,
the inner block captured in a CUDA graph, snapshots in a device ring)."""
from __future__ import annotations

from typing import Callable


def _identity(x):
    return x


def repeated(f: Callable, steps: int) -> Callable:
    adv = getattr(f, "advance", None)
    if adv is not None:
        run = lambda x: adv(x, steps)
        run._jaxib_repeated = (f, steps)  # lets `trajectory` fuse the two loops
        return run

    def run(x):
        for _ in range(steps):
            x = f(x)
        return x

    return run


def trajectory(step_fn: Callable, steps: int, post_process: Callable = _identity, start_with_input: bool = False):
    """Returns f(x) -> (final, [snapshot_0 .. snapshot_{steps-1}]) (a list instead of a stacked pytree)."""
    fused, inner = getattr(step_fn, "_jaxib_repeated", (step_fn, 1))
    roll = getattr(fused, "rollout", None)
    if roll is not None and post_process is _identity:
        def run_device(x):
            out = roll(x, inner, steps, start_with_input)
            return out if out is not None else run(x)

        def run(x):
            snaps = []
            for _ in range(steps):
                nxt = step_fn(x)
                snaps.append(x if start_with_input else nxt)
                x = nxt
            return x, snaps

        return run_device

    def run(x):
        snaps = []
        for _ in range(steps):
            nxt = step_fn(x)
            snaps.append(post_process(x if start_with_input else nxt))
            x = nxt
        return x, snaps

    return run
