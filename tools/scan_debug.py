#!/usr/bin/env python
"""Debug aid: pipelined scan forward / backward on growing sequence lengths, checked against the two-pass schedule."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import _lib, ops  # noqa: E402

d = torch.device("cuda:0")
H = int(os.environ.get("H", "32"))
Di = 16 * H
B = int(os.environ.get("B", "1"))
only_fwd = os.environ.get("FWD_ONLY", "0") == "1"
for L in [int(s) for s in (sys.argv[1] if len(sys.argv) > 1 else "2500,4096,8192,16384,65536").split(",")]:
    g = torch.Generator().manual_seed(L)
    mk = lambda *s: torch.randn(*s, generator=g).to(d, torch.bfloat16)
    xa, z, BC = mk(B, L, Di), mk(B, L, Di), mk(B, L, 2 * Di)
    dlog = (torch.randn(B, L, H, generator=g) - 3.0).to(d, torch.bfloat16)
    A_log = (torch.rand(H, 16, generator=g) * 0.68 - 0.69).to(d)
    D = torch.ones(Di, device=d)
    dy = mk(B, L, Di)
    res = {}
    for mode in (_lib.SCAN_TWO_PASS, _lib.SCAN_PIPELINED):
        leaves = [t.detach().clone().requires_grad_(True) for t in (xa, dlog, BC, z, A_log, D)]
        t0 = time.time()
        y = ops.selective_scan(*leaves, mode=mode)[0]
        torch.cuda.synchronize()
        t1 = time.time()
        if not only_fwd:
            y.backward(dy)
            torch.cuda.synchronize()
        t2 = time.time()
        res[mode] = [y.detach().float()] + ([t.grad.float() for t in leaves] if not only_fwd else [])
        print(f"L={L} mode={mode} fwd {1e3 * (t1 - t0):.1f} ms bwd {1e3 * (t2 - t1):.1f} ms", flush=True)
    names = ["y", "dxa", "ddlog", "dBC", "dz", "dA_log", "dD"]
    for n, a, b in zip(names, res[_lib.SCAN_TWO_PASS], res[_lib.SCAN_PIPELINED]):
        print(f"   {n}: rel {float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)):.2e}", flush=True)
