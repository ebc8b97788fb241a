#!/usr/bin/env python
"""Probe: torch symmetric memory (peer-mapped buffers over NVLink) under torchrun on this box.
    python -m torch.distributed.run --nproc-per-node 2 tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    g = dist.group.WORLD
    t = symm_mem.empty(world, 1 << 20, dtype=torch.float32, device=dev)
    t.zero_()
    h = symm_mem.rendezvous(t, g)
    print(rank, "rendezvous ok: ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support, flush=True)
    h.barrier(channel=0)
    # every rank writes its id into block `rank` of every peer's buffer (plain torch copy into the peer-mapped view)
    for peer in range(world):
        pv = h.get_buffer(peer, (world, 1 << 20), torch.float32)
        pv[rank].fill_(float(rank + 1))
    h.barrier(channel=1)
    torch.cuda.synchronize()
    ok = all(float(t[s, 12345]) == s + 1 for s in range(world))
    print(rank, "peer writes visible:", ok, flush=True)
    # barrier cost, eager and inside a CUDA graph
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(50):
        h.barrier(channel=0)
    ev[1].record()
    torch.cuda.synchronize()
    print(rank, "barrier us (eager):", ev[0].elapsed_time(ev[1]) * 1e3 / 50, flush=True)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(10):
                    h.barrier(channel=0)
            torch.cuda.synchronize()
            dist.barrier()
            ev[0].record(s)
            for _ in range(5):
                gr.replay()
            ev[1].record(s)
            torch.cuda.synchronize()
        print(rank, "barrier us (graph replay):", ev[0].elapsed_time(ev[1]) * 1e3 / 50, flush=True)
    except Exception as ex:
        print(rank, "graph capture of barrier failed:", repr(ex)[:300], flush=True)
    # bandwidth of a peer copy (rank 0 -> rank 1's buffer)
    if world > 1:
        src = torch.randn(64 << 20, device=dev)      # 256 MB
        big = symm_mem.empty(64 << 20, dtype=torch.float32, device=dev)
        hb = symm_mem.rendezvous(big, g)
        hb.barrier(channel=0)
        dstv = hb.get_buffer((rank + 1) % world, (64 << 20,), torch.float32)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(5):
            dstv.copy_(src)
        ev[1].record()
        torch.cuda.synchronize()
        print(rank, "peer store bandwidth GB/s:", 5 * 256e-3 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1.0, flush=True)
        ev[0].record()
        for _ in range(5):
            src.copy_(dstv)
        ev[1].record()
        torch.cuda.synchronize()
        print(rank, "peer load bandwidth GB/s:", 5 * 256e-3 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1.0, flush=True)
        hb.barrier(channel=0)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
