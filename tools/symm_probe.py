"""Probe (torchrun, >= 2 GPUs): can this box map peer GPU memory into the local address space?
Tries torch.distributed._symmetric_memory, prints peer pointers, barrier latency and peer-copy bandwidth."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    try:
        import torch.distributed._symmetric_memory as symm_mem

        n = 64 << 20
        t = symm_mem.empty(n // 4, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        if rank == 0:
            print("symm_mem ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs",
                  [hex(p) for p in hdl.signal_pad_ptrs], "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0), flush=True)
        t.fill_(float(rank))
        hdl.barrier()
        peer = hdl.get_buffer((rank + 1) % world, (n // 4,), torch.float32)
        src = torch.full((n // 4,), 100.0 + rank, device=device)
        torch.cuda.synchronize()
        for _ in range(3):
            peer.copy_(src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            peer.copy_(src)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        hdl.barrier()
        ok = bool((t == 100.0 + (rank - 1) % world).all().item())
        e0.record()
        for _ in range(100):
            hdl.barrier()
        e1.record()
        torch.cuda.synchronize()
        print(f"rank {rank}: peer write {n / ms / 1e6:.1f} GB/s ({ms * 1e3:.1f} us for 64 MiB), data ok={ok}, "
              f"barrier {e0.elapsed_time(e1) * 10:.2f} us", flush=True)
        # NCCL all_to_all for comparison
        a = torch.empty(n // 4, device=device)
        b = torch.empty(n // 4, device=device)
        for _ in range(3):
            dist.all_to_all_single(b, a)
        e0.record()
        for _ in range(10):
            dist.all_to_all_single(b, a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"rank {rank}: nccl all_to_all_single 64 MiB total: {ms * 1e3:.1f} us", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"rank {rank}: symm_mem FAILED: {type(e).__name__}: {e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
