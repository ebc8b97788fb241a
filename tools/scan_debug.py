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
    for key, ent in ops._scan_ws.items():
        hdr = ent[0][:128].view(torch.int32)[[0, 1, 2, 16]].tolist()
        print("   ws", key[2], "sync/err words:", hdr, flush=True)
        if key[2] and hdr[3]:
            # protocol error: first tile per chain (in backward scan order) without aggregate / without incoming state
            TTb, Cs = int(os.environ.get("TTB", "64")), 64
            nch_b, nch_f = -(-L // TTb), -(-L // 48)
            ntiles = B * (Di // Cs) * max(nch_b, nch_f)
            words = ent[0][128:128 + ntiles * Cs * 16].view(torch.int64).view(ntiles, Cs, 2)
            incl = ent[0][128 + ntiles * Cs * 16:128 + ntiles * Cs * 24].view(torch.int64).view(ntiles, Cs)
            ep = hdr[2]          # epoch of the last finished launch = stored value (launch used stored + 1 before increment)
            wv = ((words >> 34) == ep).all(dim=2).all(dim=1).cpu()
            iv = ((incl >> 34) == ep).all(dim=1).cpu()
            for chain in range(B * (Di // Cs)):
                a = wv[chain * nch_b:(chain + 1) * nch_b].flip(0)
                i = iv[chain * nch_b:(chain + 1) * nch_b].flip(0)
                print(f"   chain {chain}: aggregates valid {int(a.sum())}/{nch_b}, first missing (scan order) {int((~a).nonzero()[0]) if (~a).any() else -1};"
                      f" incoming valid {int(i.sum())}/{nch_b}, first missing {int((~i).nonzero()[0]) if (~i).any() else -1}", flush=True)
    names = ["y", "dxa", "ddlog", "dBC", "dz", "dA_log", "dD"]
    for n, a, b in zip(names, res[_lib.SCAN_TWO_PASS], res[_lib.SCAN_PIPELINED]):
        print(f"   {n}: rel {float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)):.2e}", flush=True)
