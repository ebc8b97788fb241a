#!/usr/bin/env python
"""Is the bench step GPU-bound or launch-bound?  Prints the host time to enqueue one step (no sync inside) next to the
device time of the same steps (CUDA events), for the bench workload."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import ApertisLayerB200, BlockConfig  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
        (out.float().pow(2).mean() + lb + rz).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print(f"batch {B}: host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, device {e0.elapsed_time(e1) / n:.2f} ms/step")
    if os.environ.get("STEP_PROFILE"):
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(5):
                step()
            torch.cuda.synchronize()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        busy = sum(e.device_time for e in ev if e.device_time) / 5
        t_lo = min(e.time_range.start for e in ev); t_hi = max(e.time_range.end for e in ev)
        print(f"  profiler: {len(ev) / 5:.0f} device activities/step, busy {busy / 1e3:.2f} ms/step, span {(t_hi - t_lo) / 5e3:.2f} ms/step")
        # largest gaps between consecutive device activities
        ev.sort(key=lambda e: e.time_range.start)
        gaps = []
        for a, b in zip(ev, ev[1:]):
            g = b.time_range.start - a.time_range.end
            if g > 5:
                gaps.append((g, a.name[:50], b.name[:50]))
        gaps.sort(reverse=True)
        print(f"  idle gaps > 5 us: {len(gaps) / 5:.0f}/step, total {sum(g for g, _, _ in gaps) / 5e3:.2f} ms/step")
        for g, a, b in gaps[:12]:
            print(f"    {g:8.1f} us after {a} -> {b}")


if __name__ == "__main__":
    main()
