#!/usr/bin/env python
"""Per-shape timing of the grouped expert GEMM (the six launches of one block step) with CUDA events.

    python tools/gemm_bench.py [--dm 704 --inter 2816 --experts 8 --rows-per-expert 5120] [--iters 20] [--only nt1,...]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apertis_llm_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dm", type=int, default=704)
    ap.add_argument("--inter", type=int, default=2816)
    ap.add_argument("--experts", type=int, default=8)
    ap.add_argument("--rows-per-expert", type=int, default=5120)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    d = torch.device("cuda:0")
    E, Dm, I, R = args.experts, args.dm, args.inter, args.rows_per_expert
    rows = E * R
    seg = torch.arange(E + 1, dtype=torch.int32, device=d) * R
    plan = dict(tile_expert=torch.arange(E, dtype=torch.int32, device=d).repeat_interleave(R // _lib.ROW_ALIGN),
                n_rows=torch.full((2,), rows, dtype=torch.int32, device=d), seg_off=seg)
    bf = lambda *s: (torch.randn(*s, device=d) * 0.1).to(torch.bfloat16)
    xn, h, hpre, dy, dh = bf(rows, Dm), bf(rows, I), bf(rows, I), bf(rows, Dm), bf(rows, I)
    w1, w2 = bf(E, I, Dm), bf(E, Dm, I)
    b1, b2 = torch.randn(E, I, device=d), torch.randn(E, Dm, device=d)
    seed = torch.tensor([12345, 678], dtype=torch.int32, device=d)
    cases = {
        "nt1_fwd_bias_gelu": lambda: ops.grouped_gemm("nt", xn, w1, plan, I, Dm, E, bias=b1, epi=_lib.EPI_BIAS_ACT, act=0, want_c2=True),
        "nt2_fwd_bias": lambda: ops.grouped_gemm("nt", h, w2, plan, Dm, I, E, bias=b2, epi=_lib.EPI_BIAS),
        "nn2_dgrad_dgelu": lambda: ops.grouped_gemm("nn", dy, w2, plan, I, Dm, E, aux=hpre, epi=_lib.EPI_DACT, act=0),
        "nn1_dgrad": lambda: ops.grouped_gemm("nn", dh, w1, plan, Dm, I, E),
        "tn2_wgrad": lambda: ops.grouped_gemm_tn(dy, h, seg, Dm, I, E),
        "tn1_wgrad": lambda: ops.grouped_gemm_tn(dh, xn, seg, I, Dm, E),
        "nt1_plain": lambda: ops.grouped_gemm("nt", xn, w1, plan, I, Dm, E),
        "nt1_fwd_bias_gelu_drop": lambda: ops.grouped_gemm("nt", xn, w1, plan, I, Dm, E, bias=b1, epi=_lib.EPI_BIAS_ACT, act=0, want_c2=True,
                                                            drop_p=0.1, drop_seed=seed),
        "nn2_dgrad_dgelu_drop": lambda: ops.grouped_gemm("nn", dy, w2, plan, I, Dm, E, aux=hpre, epi=_lib.EPI_DACT, act=0, drop_p=0.1,
                                                          drop_seed=seed),
    }
    # library reference point: the same flops as one dense bf16 GEMM through cuBLAS (torch.matmul), not part of the product
    wd1, wd2 = bf(I, Dm), bf(Dm, I)
    cases["cublas_dense_nt1"] = lambda: torch.matmul(xn, wd1.t())
    cases["cublas_dense_nt2"] = lambda: torch.matmul(h, wd2.t())
    only = [c for c in args.only.split(",") if c]
    flops = 2.0 * rows * Dm * I
    for name, fn in cases.items():
        if only and name not in only:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        print(f"{name:22s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
