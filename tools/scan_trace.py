#!/usr/bin/env python
"""Phase timeline of the single-pass forward scan (debug builds only: `make -C apertis_llm_b200/csrc TRACE=1`).

Reads the per-tile %globaltimer marks the kernel leaves in g_scan_trace and prints the mean duration of each phase.
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import _lib, ops  # noqa: E402

PHASES_BWD = ["stage dt + hstart + wait TMA", "A: g, run aggregates + sync", "B: tile aggregate, publish", "C: recompute + dxa/dC/dz",
              "D: wait incoming (thread 0)", "sync", "E: reverse sweep + partials"]
PHASES = ["stage dt + issue TMA", "wait TMA", "sweep 1", "aggregate loop (thread 0)", "wait incoming", "prefix + sync", "sweep 2"]


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    mode = _lib.SCAN_TWO_PASS if len(sys.argv) > 2 and sys.argv[2] == "two_pass" else _lib.SCAN_SINGLE_PASS
    d = torch.device("cuda:0")
    H, B = 32, 1
    Di = 16 * H
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: torch.randn(*s, generator=g).to(d, torch.bfloat16)
    xa, z, BC = mk(B, L, Di), mk(B, L, Di), mk(B, L, 2 * Di)
    dlog = (torch.randn(B, L, H, generator=g) - 3.0).to(d, torch.bfloat16)
    A_log = (torch.rand(H, 16, generator=g) * 0.68 - 0.69).to(d)
    D = torch.ones(Di, device=d)
    lib = _lib.load()
    sbuf = np.zeros((64, 2048, 4), dtype=np.uint64)
    lib.ab_scanner_trace_dump.argtypes = [ctypes.c_void_p, ctypes.c_int]
    bwd = os.environ.get("TRACE_BWD")
    if bwd:
        for t in (xa, dlog, BC, z):
            t.requires_grad_(True)
        dy = mk(B, L, Di)
    for _ in range(3):
        lib.ab_scanner_trace_dump(sbuf.ctypes.data, 1)
        y = ops.selective_scan(xa, dlog, BC, z, A_log, D, mode=mode)[0]
        if bwd:
            torch.cuda.synchronize()
            lib.ab_scanner_trace_dump(sbuf.ctypes.data, 1)
            y.backward(dy)
    torch.cuda.synchronize()
    lib.ab_scanner_trace_dump(sbuf.ctypes.data, 0)
    r = sbuf[0]
    r = r[r[:, 0] > 0].astype(np.int64)
    if len(r):
        print(f"scanner chain 0: {len(r)} rounds, tiles/round mean {r[:, 3].mean():.1f}; load wait mean {(r[:, 1] - r[:, 0]).mean() / 1e3:.2f} us, "
              f"consume mean {(r[:, 2] - r[:, 1]).mean() / 1e3:.2f} us, round period mean {np.diff(r[:, 0]).mean() / 1e3:.2f} us")
        print("   tiles/round histogram:", np.bincount(r[:, 3].astype(int), minlength=33).tolist())
    n = 16384
    buf = np.zeros((n, 8), dtype=np.uint64)
    lib.ab_scan_trace_dump.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = lib.ab_scan_trace_dump(buf.ctypes.data, n)
    assert rc == 0, rc
    live = buf[:, 0] > 0
    t = buf[live].astype(np.int64)
    print(f"tiles traced: {live.sum()}, kernel span {(t[:, 7].max() - t[:, 0].min()) / 1e3:.1f} us")
    life = (t[:, 7] - t[:, 0])
    print(f"tile lifetime mean {life.mean() / 1e3:.2f} us  p50 {np.median(life) / 1e3:.2f}  p95 {np.percentile(life, 95) / 1e3:.2f}")
    for k, name in enumerate(PHASES_BWD if bwd else PHASES):
        dt = t[:, k + 1] - t[:, k]
        ok = (t[:, k + 1] > 0) & (t[:, k] > 0)
        if ok.any():
            print(f"  {name:28s} mean {dt[ok].mean() / 1e3:7.2f} us  p50 {np.median(dt[ok]) / 1e3:7.2f}  p95 {np.percentile(dt[ok], 95) / 1e3:7.2f}")
    # concurrency: tiles alive at the same time
    ev = np.concatenate([np.stack([t[:, 0], np.ones(len(t), np.int64)], 1), np.stack([t[:, 7], -np.ones(len(t), np.int64)], 1)])
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    print(f"max tiles alive: {np.cumsum(ev[:, 1]).max()}")
    if mode == _lib.SCAN_SINGLE_PASS and os.environ.get("TRACE_DETAIL"):
        chain_detail(buf, 8192 // 8 if L == 65536 else int(os.environ["TRACE_NCHUNKS"]))



def chain_detail(t_all, nchunks, chain=0, j0=500, n=48):
    base = t_all[:, 0][t_all[:, 0] > 0].min()
    print("  j   start  published  received  end   (us since kernel start)")
    for j in range(j0, j0 + n):
        r = t_all[chain * nchunks + j].astype(np.int64)
        print(f"{j:5d} {(r[0] - base) / 1e3:8.2f} {(r[4] - base) / 1e3:8.2f} {(r[5] - base) / 1e3:8.2f} {(r[7] - base) / 1e3:8.2f}")


if __name__ == "__main__":
    main()
