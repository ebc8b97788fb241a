"""The "rounds" selective scan (csrc/ssm_scan_rounds.cu, the default schedule) on the B200, through the C ABI, against the
CPU oracle's recurrent scan (core.py:337-353) - the reference's finite formulation at every length (SURVEY.md 8c caveat 1:
its training-mode log-cumsum form returns NaN from L = 16K)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import apertis_oracle as O
from tests.util import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def dev():
    return torch.device("cuda:0")


def make_case(B, L, H, dtype, seed, h0=True):
    Di = 16 * H
    g = torch.Generator().manual_seed(seed)
    q = lambda t: t.to(dtype).float()
    c = dict(xa=q(torch.randn(B, L, Di, generator=g)), z=q(torch.randn(B, L, Di, generator=g)),
             BC=q(torch.randn(B, L, 2 * Di, generator=g) * 0.5), dlog=q(torch.randn(B, L, H, generator=g) - 3.0),
             A_log=torch.rand(H, 16, generator=g) * (math.log(0.99) - math.log(0.5)) + math.log(0.5),
             D=1.0 + 0.1 * torch.randn(Di, generator=g), h0=torch.randn(B, H, 16, generator=g) if h0 else None,
             dy=q(torch.randn(B, L, Di, generator=g)), dys=q(torch.randn(B, L, Di, generator=g) * 0.3))
    return c


def oracle_scan(c, heads=None, chunk=8192, want_yssm=True):
    """Outputs and gradients of the scan + skip + gate from O.scan_recurrent in float64, optionally for a subset of heads
    (heads are independent chains), with the time loop cut into chunks whose boundary state / its gradient are handed on
    explicitly (same arithmetic; keeps autograd graphs short for the 16K-64K cases)."""
    B, L, Di = c["xa"].shape
    H = c["dlog"].shape[-1]
    hs = list(range(H)) if heads is None else list(heads)
    ch = torch.tensor([h * 16 + n for h in hs for n in range(16)])
    Hs = len(hs)
    sel = lambda t: t[..., ch].double()
    xa, z, dy, dys = sel(c["xa"]), sel(c["z"]), sel(c["dy"]), sel(c["dys"])
    Bm, Cm = sel(c["BC"][..., :Di]), sel(c["BC"][..., Di:])
    dlog = c["dlog"][..., hs].double()
    A_log, D = c["A_log"][hs].double(), c["D"][ch].double()
    h0 = c["h0"][:, hs].double() if c["h0"] is not None else torch.zeros(B, Hs, 16, dtype=torch.float64)
    bounds = list(range(0, L, chunk)) + [L]
    # forward, chunk by chunk, keeping the boundary states
    states, ys_all, y_all = [h0], [], []
    with torch.no_grad():
        for a, b in zip(bounds[:-1], bounds[1:]):
            delta = F.softplus(dlog[:, a:b]).transpose(1, 2).unsqueeze(-1)
            ys, hl = O.scan_recurrent(delta, A_log, Bm[:, a:b].reshape(B, b - a, Hs, 16).transpose(1, 2),
                                      Cm[:, a:b].reshape(B, b - a, Hs, 16).transpose(1, 2), states[-1])
            ys = ys.transpose(1, 2).reshape(B, b - a, Hs * 16)
            ys_all.append(ys)
            y_all.append((ys + D * xa[:, a:b]) * F.silu(z[:, a:b]))
            states.append(hl)
    out = dict(y=torch.cat(y_all, 1), y_ssm=torch.cat(ys_all, 1), h_last=states[-1].reshape(B, Hs * 16), ch=ch, hs=hs)
    # backward, last chunk first, carrying d(state)
    g = {k: torch.zeros_like(v) for k, v in dict(xa=xa, z=z, Bm=Bm, Cm=Cm, dlog=dlog, A_log=A_log, D=D).items()}
    dh = torch.zeros_like(h0)
    for i in range(len(bounds) - 2, -1, -1):
        a, b = bounds[i], bounds[i + 1]
        lv = dict(xa=xa[:, a:b], z=z[:, a:b], Bm=Bm[:, a:b], Cm=Cm[:, a:b], dlog=dlog[:, a:b], A_log=A_log, D=D, h=states[i])
        lv = {k: v.detach().clone().requires_grad_(True) for k, v in lv.items()}
        delta = F.softplus(lv["dlog"]).transpose(1, 2).unsqueeze(-1)
        ys, hl = O.scan_recurrent(delta, lv["A_log"], lv["Bm"].reshape(B, b - a, Hs, 16).transpose(1, 2),
                                  lv["Cm"].reshape(B, b - a, Hs, 16).transpose(1, 2), lv["h"])
        ys = ys.transpose(1, 2).reshape(B, b - a, Hs * 16)
        y = (ys + lv["D"] * lv["xa"]) * F.silu(lv["z"])
        loss = (y * dy[:, a:b]).sum() + (hl * dh).sum()
        if want_yssm:
            loss = loss + (ys * dys[:, a:b]).sum()
        loss.backward()
        for k in ("xa", "z", "Bm", "Cm", "dlog"):
            g[k][:, a:b] = lv[k].grad
        g["A_log"] += lv["A_log"].grad
        g["D"] += lv["D"].grad
        dh = lv["h"].grad
    out["grads"] = g
    return out


def run_gpu(c, dtype, want_yssm=True, fused=False, strided=False, dt_bias=None):
    from apertis_llm_b200 import ops
    B, L, Di = c["xa"].shape
    H = c["dlog"].shape[-1]
    d = dev()
    A_log = c["A_log"].to(d).requires_grad_(True)
    D = c["D"].to(d).requires_grad_(True)
    h0 = c["h0"].to(d) if c["h0"] is not None else None
    if strided:       # xa and z as column slices of one wider buffer, the way the fused in-projection output is consumed
        xz = torch.cat([c["xa"], c["z"]], dim=-1).to(d, dtype).requires_grad_(True)
        xa, z = xz[..., :Di], xz[..., Di:]
    else:
        xa, z = c["xa"].to(d, dtype).requires_grad_(True), c["z"].to(d, dtype).requires_grad_(True)
    bias = dt_bias.to(d).requires_grad_(True) if dt_bias is not None else None
    dl_in = c["dlog"] - (dt_bias if dt_bias is not None else 0.0)
    if fused:
        Hp = (H + 7) // 8 * 8
        prm = torch.cat([dl_in, torch.full((B, L, Hp - H), 7.0), c["BC"]], dim=-1).to(d, dtype).requires_grad_(True)
        y, ys, hl = ops.selective_scan_fused(xa, prm, bias, z, A_log, D, H, h0=h0, want_yssm=want_yssm, want_hlast=True)
    else:
        dlog = dl_in.to(d, dtype).requires_grad_(True)
        BC = c["BC"].to(d, dtype).requires_grad_(True)
        y, ys, hl = ops._SelectiveScanRounds.apply(xa, dlog, bias, BC, z, A_log, D, h0, want_yssm, True, H)
    outs, gos = [y], [c["dy"].to(d, dtype)]
    if want_yssm:
        outs.append(ys); gos.append(c["dys"].to(d, dtype))
    torch.autograd.backward(outs, gos)
    torch.cuda.synchronize()
    r = dict(y=y.detach().float().cpu(), y_ssm=ys.detach().float().cpu() if want_yssm else None, h_last=hl.cpu(),
             dA_log=A_log.grad.cpu(), dD=D.grad.cpu())
    if strided:
        r["dxa"], r["dz"] = xz.grad[..., :Di].float().cpu(), xz.grad[..., Di:].float().cpu()
    else:
        r["dxa"], r["dz"] = xa.grad.float().cpu(), z.grad.float().cpu()
    if fused:
        Hp = prm.shape[-1] - 2 * Di
        gp = prm.grad.float().cpu()
        r["ddlog"], r["dpad"], r["dBC"] = gp[..., :H], gp[..., H:Hp], gp[..., Hp:]
    else:
        r["ddlog"], r["dBC"] = dlog.grad.float().cpu(), BC.grad.float().cpu()
    if bias is not None:
        r["dbias"] = bias.grad.cpu()
    return r


def compare(r, o, dtype, want_yssm=True, loose=()):
    tol = TOL[dtype]
    ch, hs = o["ch"], o["hs"]
    Di = r["y"].shape[-1]
    assert rel_err(r["y"][..., ch], o["y"]) < tol, "y"
    if want_yssm:
        assert rel_err(r["y_ssm"][..., ch], o["y_ssm"]) < tol, "y_ssm"
    assert rel_err(r["h_last"][..., ch], o["h_last"]) < tol, "h_last"
    g = o["grads"]
    red = max(tol, 2e-3) if dtype == torch.bfloat16 else tol         # reductions over all tokens of bf16-rounded terms
    checks = [("dxa", r["dxa"][..., ch], g["xa"], tol), ("dz", r["dz"][..., ch], g["z"], tol),
              ("dB", r["dBC"][..., :Di][..., ch], g["Bm"], tol), ("dC", r["dBC"][..., Di:][..., ch], g["Cm"], tol),
              ("ddlog", r["ddlog"][..., hs], g["dlog"], red), ("dA_log", r["dA_log"][hs], g["A_log"], red),
              ("dD", r["dD"][ch], g["D"], red)]
    for name, a, b, t in checks:
        assert rel_err(a, b) < t, name


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,L,H", [(2, 48, 2), (1, 1000, 11), (2, 333, 4), (1, 5, 3), (1, 2500, 32), (3, 777, 8), (1, 1, 4), (2, 8, 25)])
def test_rounds_scan_vs_oracle(dtype, B, L, H):
    c = make_case(B, L, H, dtype, seed=L + H)
    o = oracle_scan(c)
    compare(run_gpu(c, dtype), o, dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rounds_scan_fused_strided_layout(dtype):
    """[dt (no bias) | pad | B | C] as one buffer, xa / z as slices of one buffer, dt bias applied inside the kernel:
    the layouts the fused projections produce.  The padding columns must come back as zero gradients."""
    B, L, H = 2, 700, 11
    c = make_case(B, L, H, dtype, seed=3)
    bias = torch.rand(H) * 2 - 5
    o = oracle_scan(c)
    r = run_gpu(c, dtype, fused=True, strided=True, dt_bias=bias)
    compare(r, o, dtype)
    assert float(r["dpad"].abs().max()) == 0.0
    assert rel_err(r["dbias"], o["grads"]["dlog"].sum((0, 1))) < (2e-3 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("tc,wps,nst", [(8, 8, 2), (16, 8, 3), (24, 16, 2), (64, 8, 3), (32, 4, 3)])
def test_rounds_scan_many_rounds(tc, wps, nst):
    """Short chunks and few resident warps force many rounds (pipelined P1 / P2, both prefix levels, the carry chain) at a
    size the oracle covers in seconds; results must not depend on the chunking beyond fp32 rounding."""
    from apertis_llm_b200 import _lib
    lib = _lib.load()
    c = make_case(2, 6000, 8, torch.float32, seed=17)
    o = oracle_scan(c)
    try:
        lib.ab_ssm_scan_tune(tc, tc, wps, nst, nst)
        r = run_gpu(c, torch.float32)
        r2 = run_gpu(c, torch.float32)
    finally:
        lib.ab_ssm_scan_tune(0, 0, 0, 0, 0)
    compare(r, o, torch.float32)
    for k in ("y", "dxa", "ddlog", "dBC", "dz", "dA_log", "dD", "h_last"):
        assert torch.equal(r[k], r2[k]), f"{k}: not bitwise repeatable"


@pytest.mark.parametrize("dtype,B,L,H,heads", [
    (torch.bfloat16, 8, 4096, 11, None),              # configs[1]: the 1.5B block's scan at the bench batch
    (torch.float32, 2, 4673, 11, None),               # configs[2]: 4096 text + 577 image tokens (odd length)
    (torch.bfloat16, 2, 4096, 25, (0, 7, 24)),        # configs[3]: 7B-class width (d_inner 400: 6 slabs + 16 channels)
    (torch.bfloat16, 1, 16384, 32, (0, 13, 31)),      # configs[4]: long-context sweep, d_inner 512
    (torch.float32, 1, 16384, 32, (5, 30)),
    (torch.bfloat16, 1, 65536, 32, (0, 17, 31)),
])
def test_rounds_scan_baseline_shapes(dtype, B, L, H, heads):
    """The scan at the BASELINE.json shapes against the recurrent oracle (a subset of heads where the CPU loop would take
    minutes: heads are independent chains, the kernel still runs the full width)."""
    c = make_case(B, L, H, dtype, seed=L + 3 * H, h0=False)
    o = oracle_scan(c, heads=heads, want_yssm=False)
    r = run_gpu(c, dtype, want_yssm=False)
    compare(r, o, dtype, want_yssm=False)
    assert torch.isfinite(r["y"]).all() and torch.isfinite(r["dxa"]).all()


def test_rounds_scan_cuda_graph_and_determinism():
    from apertis_llm_b200 import ops
    B, L, H = 2, 6000, 8
    Di = 16 * H
    g = torch.Generator().manual_seed(11)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev(), torch.bfloat16)
    xa, z, BC, dy = mk(B, L, Di), mk(B, L, Di), mk(B, L, 2 * Di) * 0.5, mk(B, L, Di)
    dlog = (torch.randn(B, L, H, generator=g) - 3).to(dev(), torch.bfloat16)
    A_log = (torch.rand(H, 16, generator=g) * 0.6 - 0.7).to(dev())
    D = torch.ones(Di, device=dev())
    leaves = [t.requires_grad_(True) for t in (xa, dlog, BC, z, A_log, D)]

    def step():
        y = ops.selective_scan(*leaves)[0]
        return [y] + list(torch.autograd.grad(y, leaves, dy))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                                # allocates this stream's workspace before the capture
        ref = [t.clone() for t in step()]
        again = step()
        torch.cuda.synchronize()
        for a, b in zip(again, ref):
            assert torch.equal(a, b), "the rounds scan is not bitwise repeatable"
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            outs = step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        for o in outs:
            o.zero_()
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(outs, ref):
            assert torch.equal(a, b), "graph replay of the rounds scan differs from the eager launch"


def test_rounds_scan_on_non_current_device_guard():
    """Entry points launch on the tensors' device, not on whatever device happens to be current (single-GPU box: the
    guard is exercised with the current device set explicitly)."""
    from apertis_llm_b200 import ops
    c = make_case(1, 64, 2, torch.float32, seed=1)
    with torch.cuda.device(0):
        r = run_gpu(c, torch.float32)
    compare(r, oracle_scan(c), torch.float32)
