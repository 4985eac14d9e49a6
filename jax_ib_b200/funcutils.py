"""`repeated` / `trajectory` with the semantics of jax_cfd.base.funcutils (un-vendored upstream the
reference's notebooks call, Flapping_Demo.ipynb:225-247).  A fused step function produced by
equations.semi_implicit_navier_stokes_timeBC exposes `.advance(x, n)`; `repeated` then runs the whole
inner loop on the GPU with one host round trip instead of n."""
from __future__ import annotations

from typing import Callable


def repeated(f: Callable, steps: int) -> Callable:
    adv = getattr(f, "advance", None)
    if adv is not None:
        return lambda x: adv(x, steps)

    def run(x):
        for _ in range(steps):
            x = f(x)
        return x

    return run


def trajectory(step_fn: Callable, steps: int, post_process: Callable = lambda x: x, start_with_input: bool = False):
    """Returns f(x) -> (final, [snapshot_0 .. snapshot_{steps-1}]) (a list instead of a stacked pytree)."""

    def run(x):
        snaps = []
        for _ in range(steps):
            nxt = step_fn(x)
            snaps.append(post_process(x if start_with_input else nxt))
            x = nxt
        return x, snaps

    return run
