"""Multi-GPU check of the slab decomposition (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/slab_check.py [--n 1024] [--steps 6]

Every rank advances its slab; the gathered fields are compared with the single-GPU step (run on every
rank's own GPU): bit for bit on the transposing paths (NCCL, or --peer with JAXIB_SLAB_FFT=1: the same
column FFT as one GPU with JAXIB_COL_SWEEP=0), rel-L2 <= 5e-6 (pressure 5e-5) on the default peer-memory
path, whose column sweep associates the carried values per slab.  Prints one `SLAB_CHECK OK ...` line on
rank 0, exits non-zero on any mismatch."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--peer", action="store_true", help="peer-memory mode (NVLink stores from the kernels)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from jax_ib_b200 import configs, demo, slab
    from test_slab import _column_fft, _seam_cfg, _single_gpu

    cfg = _seam_cfg(args.n)
    bitwise = not args.peer or os.environ.get("JAXIB_SLAB_FFT") == "1"
    if bitwise:
        with _column_fft():
            ref = _single_gpu(cfg, args.steps)
    else:
        ref = _single_gpu(cfg, args.steps)
    st = slab.SlabStepper(demo.make_spec(cfg), demo.make_particles(cfg), device=device, peer=args.peer)
    st.load(*(torch.from_numpy(a) for a in configs.initial_fields(cfg)))
    st.advance(2)
    st.advance(args.steps - 2)
    got = st.gather()
    timed_out = args.peer and st.ranks[0].barrier_timed_out()
    worst = 0.0
    ok = True
    for name, a, b in zip("uvp", got, ref):
        same = bool(torch.equal(a, b)) if bitwise else float((a - b).norm() / b.norm()) < (5e-5 if name == "p" else 5e-6)
        ok = ok and same and bool(torch.isfinite(a).all())
        worst = max(worst, float((a - b).abs().max()))
    flag = torch.tensor([0 if ok and not timed_out else 1], device=device)
    if world > 1:
        dist.all_reduce(flag)
    if rank == 0:
        print(f"SLAB_CHECK {'OK' if flag.item() == 0 else 'FAIL'} world={world} n={args.n} steps={args.steps} "
              f"max_abs_diff={worst:.3e} compare={'bitwise' if bitwise else 'rel-L2'} mode={'peer-memory' if args.peer else type(st.comm).__name__}", flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
