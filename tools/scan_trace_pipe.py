#!/usr/bin/env python
"""Phase timeline of the pipelined forward scan (debug build: make -C apertis_llm_b200/csrc TRACE=1).

Marks per warp and iteration: 0 top | 1 incoming state resolved | 2 main operands landed | 3 main pass done |
4 prepass operands landed | 5 prepass done | 6 past the barrier | 7 end (refill, duty, coefficient queue)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import _lib, ops  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = torch.device("cuda:0")
H, B = 32, 1
Di = 16 * H
g = torch.Generator().manual_seed(L)
mk = lambda *s: torch.randn(*s, generator=g).to(d, torch.bfloat16)
xa, z, BC = mk(B, L, Di), mk(B, L, Di), mk(B, L, 2 * Di)
dlog = (torch.randn(B, L, H, generator=g) - 3.0).to(d, torch.bfloat16)
A_log = (torch.rand(H, 16, generator=g) * 0.68 - 0.69).to(d)
D = torch.ones(Di, device=d)
lib = _lib.load()
CT, IT, W, S = 8, 48, 16, 8
buf = np.zeros(CT * IT * W * S, dtype=np.uint64)
for rep in range(3):
    lib.ab_pipe_trace_dump(buf.ctypes.data_as(ctypes.c_void_p), 1)
    ops.selective_scan(xa, dlog, BC, z, A_log, D, mode=_lib.SCAN_PIPELINED)
    torch.cuda.synchronize()
lib.ab_pipe_trace_dump(buf.ctypes.data_as(ctypes.c_void_p), 0)
t = buf.reshape(CT, IT, W, S).astype(np.int64)
names = ["poll", "main wait", "main pass", "pre wait", "pre pass", "barrier", "tail", "loop back"]
for cta in range(CT):
    x = t[cta]
    ok = (x[:, :, 0] > 0) & (x[:, :, 7] > 0)
    its = [i for i in range(8, IT - 1) if ok[i].all() and ok[i + 1].all()]
    if not its:
        continue
    seg = np.zeros((len(its), W, 8))
    for n, i in enumerate(its):
        for k in range(7):
            a, b = x[i, :, k], x[i, :, k + 1]
            seg[n, :, k] = np.where((a > 0) & (b > 0), b - a, 0)
        seg[n, :, 7] = x[i + 1, :, 0] - x[i, :, 7]
    period = np.mean([x[i + 1, 0, 0] - x[i, 0, 0] for i in its])
    print(f"CTA {cta}: {len(its)} steady iterations, period {period:.0f} ns; mean over warps [ns]: " +
          " | ".join(f"{nm} {seg[:, :, k].mean():.0f}" for k, nm in enumerate(names)))
    print("      max over warps: " + " | ".join(f"{nm} {seg[:, :, k].max(axis=1).mean():.0f}" for k, nm in enumerate(names)))
