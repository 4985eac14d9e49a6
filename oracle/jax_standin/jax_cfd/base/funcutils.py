"""jax_cfd.base.funcutils.repeated / trajectory as the notebooks use them (Flapping_Demo.ipynb cell 9).  The originals
drive the step with jax.lax.scan, which turns every Python-scalar LEAF of the carried pytree into a 32-bit array before
the first step (x64 is off) -- notably the clock `time_stamp=0.0` the notebooks put into the velocity BC, which then
accumulates dt in float32.  `_as_carry` reproduces that conversion; the loop itself is a plain Python loop."""
import numpy as _np

from jax import tree_util as _tu


def _as_carry(x):
    def leaf(v):
        if isinstance(v, bool):
            return v
        if isinstance(v, float):
            return _np.float32(v)
        if isinstance(v, int):
            return _np.int32(v)
        return v
    return _tu.tree_map(leaf, x)


def repeated(f, steps):
    def g(x):
        x = _as_carry(x)
        for _ in range(steps):
            x = f(x)
        return x
    return g


def trajectory(step_fn, steps, post_process=lambda x: x, start_with_input=False):
    """(final state, snapshots stacked leaf by leaf along a new leading axis); with start_with_input the snapshot is
    taken BEFORE each step."""
    def run(x):
        x = _as_carry(x)
        snaps = []
        for _ in range(steps):
            nxt = step_fn(x)
            snaps.append(post_process(x if start_with_input else nxt))
            x = nxt
        flat = [_tu.tree_flatten(s) for s in snaps]
        stacked = [_np.stack([_np.asarray(f[0][k]) for f in flat]) for k in range(len(flat[0][0]))]
        return x, _tu.tree_unflatten(flat[0][1], stacked)
    return run
