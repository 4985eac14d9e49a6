"""Per-kernel durations of the peer-memory slab step, all slabs as virtual slabs of ONE GPU in lockstep (CUPTI via
torch.profiler; run with JAXIB_NO_PDL=1 so that a kernel's record does not include waiting for its predecessor).

    JAXIB_NO_PDL=1 python tools/slab_kernel_times.py [N] [WORLD]
Prints, per kernel name, launches per step and the mean / max duration over the slabs: what one slab of a real
multi-GPU run spends in each kernel (without the NVLink stores and the flag barriers)."""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    from jax_ib_b200 import configs, demo, slab
    cfg = configs.ib_periodic(n, max(1, n // 1024))
    st = slab.SlabStepper(demo.make_spec(cfg), demo.make_particles(cfg), world=world, virtual=True, peer=True)
    st.load(*(torch.from_numpy(a) for a in configs.initial_fields(cfg)))
    st.advance(3)
    torch.cuda.synchronize()
    steps = 3
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        st.advance(steps)
        torch.cuda.synchronize()
    acc = collections.OrderedDict()
    for e in sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start):
        if "Memcpy" in e.name or "Memset" in e.name:
            continue
        m = re.search(r"(\w+_kernel)(<[^>]*>)?", e.name)
        name = (m.group(0) if m else e.name)[:48]
        acc.setdefault(name, []).append(e.time_range.end - e.time_range.start)
    total = 0.0
    print(f"slab kernels, {n}^2 over {world} virtual slabs, {steps} steps:")
    for name, ds in acc.items():
        per_step_per_slab = len(ds) / steps / world
        mean = sum(ds) / len(ds)
        total += mean * per_step_per_slab
        print(f"  {name:48s} {per_step_per_slab:5.2f} launches/slab/step  mean {mean:7.1f} us  max {max(ds):7.1f} us")
    print(f"  sum of means per slab and step: {total:.1f} us")


if __name__ == "__main__":
    main()
