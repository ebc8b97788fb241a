#!/usr/bin/env python
"""Per-kernel device time of one eager block step (bench workload) from torch.profiler: where the non-GEMM time goes.

    python tools/step_kernels.py [batch] [top]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import ApertisLayerB200, BlockConfig  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    dev = torch.device("cuda:0")
    cfg = BlockConfig(hidden_size=704, num_attention_heads=11, intermediate_size=2816, num_experts=8, experts_per_token=2,
                      hidden_dropout_prob=0.1)
    layer = ApertisLayerB200(cfg).to(dev).train()
    x = torch.randn(B, 4096, 704, device=dev, requires_grad=True)

    def step():
        for p in layer.parameters():
            p.grad = None
        x.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, _, _, lb, rz = layer(x)
        n = torch.linalg.vector_norm(out, 2, dtype=torch.float32)
        (n * n / out.numel() + lb + rz).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / n, e.count / n) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"device time per step {tot:.0f} us over {sum(r[2] for r in rows):.0f} kernels")
    for k, t, c in rows[:top]:
        print(f"{t:8.1f} us {c:5.1f}x {100 * t / tot:5.1f}%  {k[:110]}")


if __name__ == "__main__":
    main()
