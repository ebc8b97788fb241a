#!/usr/bin/env python
"""Short digest of a bench.py JSON line:  python tools/bench_brief.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(f"value {d.get('value', 0) / 1e6:.2f} M tok/s  {d.get('ms_per_step', 0):.3f} ms/step  e2e {d.get('e2e', {}).get('value', 0) / 1e6:.2f} M  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
print(f"  gemm {r.get('achieved', 0):.0f} TF/s frac {r.get('frac', 0):.3f} ({r.get('ms_per_step_in_kernel', 0):.3f} ms, share {r.get('share_of_step', 0):.2f})")
for k, v in d.get("kernels", {}).items():
    if "error" in v:
        print("  ", k, v)
        continue
    print(f"  {k}: {v.get('achieved', 0):.0f} {v.get('unit')} frac {v.get('frac', 0):.3f} {v.get('ms_per_step_in_kernel', '')} {v.get('fwd_us', '')} {v.get('bwd_us', '')}")
if "entry_point_us" in d:
    print("   per call (us, last eager step):", {k: v for k, v in d["entry_point_us"].items() if "gemm" in k})
for n, w in d.get("workloads", {}).items():
    if "error" in w:
        print("  ", n, w)
        continue
    if "entry_point_us" in w:
        print("   per call (us):", {k: v for k, v in w["entry_point_us"].items() if "grouped" in k})
    print(f"  [{n}] {w['value'] / 1e6:.2f} M tok/s {w['ms_per_step']:.3f} ms  gemm {w['roofline']['achieved']:.0f} TF/s frac {w['roofline']['frac']:.3f}  scan {w['kernels']['selective_scan_fwd+bwd']['frac']:.3f}")
for k in ("ep_parity", "gpu_reference", "cpu_baseline"):
    if k in d:
        print("  ", k, json.dumps(d[k])[:400])
