#!/usr/bin/env python
"""Where does an expert-parallel step spend its device time?  torch.profiler on rank 0 over a few eager steps of the bench
workload (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ep_profile.py"""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import ApertisLayerB200, BlockConfig  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = BlockConfig(hidden_size=704, num_attention_heads=11, intermediate_size=2816, num_experts=8, experts_per_token=2,
                      hidden_dropout_prob=0.1)
    torch.manual_seed(0)
    layer = ApertisLayerB200(cfg, ep_group=dist.group.WORLD).to(dev).train()
    replicated = [p for n, p in layer.named_parameters() if ".expert_" not in n]
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x = torch.randn(8, 4096, 704, generator=g).to(dev).requires_grad_(True)

    def step():
        for p in layer.parameters():
            p.grad = None
        x.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, _, _, lb, rz = layer(x)
        (out.float().pow(2).mean() + lb + rz).backward()
        flat = torch.cat([p.grad.reshape(-1) for p in replicated])
        dist.all_reduce(flat)

    for _ in range(5):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / n * 1e3
    dist.barrier(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        rows = [(e.key, e.device_time_total / n, e.count / n) for e in prof.key_averages() if e.device_time_total > 0]
        rows.sort(key=lambda r: -r[1])
        tot = sum(r[1] for r in rows)
        nccl = sum(r[1] for r in rows if "nccl" in r[0].lower())
        print(f"world {world}: eager step {wall:.0f} us on the stream clock; kernel time {tot:.0f} us of which NCCL {nccl:.0f} us")
        for k, t, c in rows[:44]:
            print(f"{t:8.1f} us {c:5.1f}x {100 * t / tot:5.1f}%  {k[:100]}")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
