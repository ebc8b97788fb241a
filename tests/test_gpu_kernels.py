"""Kernel-level parity on the B200, through the C ABI (ctypes), against the CPU oracle / plain torch fp32."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import apertis_oracle as O
from tests.util import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}     # north_star tolerances (max|a-b| / max|b|)


def dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------
# causal conv1d + SiLU
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,L,Di", [(2, 37, 32), (1, 300, 176), (3, 16, 64), (2, 1, 32), (1, 129, 512)])
def test_conv_silu(dtype, B, L, Di):
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + L)
    xp = torch.randn(B, L, Di, generator=g)
    w = torch.rand(Di, 1, 4, generator=g) - 0.5
    b = torch.rand(Di, generator=g) - 0.5
    dy = torch.randn(B, L, Di, generator=g)
    xr = xp.to(dtype).float().detach().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.silu(F.conv1d(xr.transpose(1, 2), wr, br, padding=3, groups=Di)[:, :, :L].transpose(1, 2))
    ref.backward(dy.to(dtype).float())
    xg = xp.to(dev(), dtype).requires_grad_(True)
    wg, bg = w.to(dev()).requires_grad_(True), b.to(dev()).requires_grad_(True)
    out = ops.causal_conv1d_silu(xg, wg, bg)
    out.backward(dy.to(dev(), dtype))
    tol = TOL[dtype]
    assert rel_err(out.float(), ref.detach()) < tol
    assert rel_err(xg.grad.float(), xr.grad) < tol
    assert rel_err(wg.grad, wr.grad) < tol and rel_err(bg.grad, br.grad) < tol


# ------------------------------------------------------------------------------------------------
# selective scan
# ------------------------------------------------------------------------------------------------
def _scan_ref(xa, dlog, BC, z, A_log, D, h0):
    B, L, Di = xa.shape
    H = dlog.shape[-1]
    delta = F.softplus(dlog).transpose(1, 2).unsqueeze(-1)
    Bt = BC[..., :Di].view(B, L, H, 16).transpose(1, 2)
    Ct = BC[..., Di:].view(B, L, H, 16).transpose(1, 2)
    y, hl = O.scan_recurrent(delta, A_log, Bt, Ct, h0)
    y_ssm = y.transpose(1, 2).reshape(B, L, Di)
    return (y_ssm + D * xa) * F.silu(z), y_ssm, hl.reshape(B, Di)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,L,H", [(2, 48, 2), (1, 1000, 11), (2, 333, 4), (1, 5, 3), (1, 2500, 32), (3, 777, 8)])
def test_selective_scan(mode, dtype, B, L, H):
    from apertis_llm_b200 import ops
    Di = 16 * H
    g = torch.Generator().manual_seed(L + H)
    q = lambda t: t.to(dtype).float()
    xa, z = q(torch.randn(B, L, Di, generator=g)), q(torch.randn(B, L, Di, generator=g))
    BC = q(torch.randn(B, L, 2 * Di, generator=g) * 0.5)
    dlog = q(torch.randn(B, L, H, generator=g) - 3.0)          # softplus -> delta ~ 0.05, decays matter over ~100 steps
    A_log = torch.rand(H, 16, generator=g) * (math.log(0.99) - math.log(0.5)) + math.log(0.5)
    D = 1.0 + 0.1 * torch.randn(Di, generator=g)
    h0 = torch.randn(B, H, 16, generator=g)
    dy, dys = q(torch.randn(B, L, Di, generator=g)), q(torch.randn(B, L, Di, generator=g) * 0.3)
    leaves = [t.clone().double().requires_grad_(True) for t in (xa, dlog, BC, z, A_log, D)]
    y, ys, hl = _scan_ref(*leaves, h0.double())
    (y * dy.double()).sum().add((ys * dys.double()).sum()).backward()
    gl = [t.to(dev(), dtype if i < 4 else torch.float32).requires_grad_(True) for i, t in enumerate((xa, dlog, BC, z, A_log, D))]
    pipelined = mode == 2            # the pipelined schedule has no y_ssm output (output_attentions falls back)
    if pipelined:
        leaves = [t.clone().double().requires_grad_(True) for t in (xa, dlog, BC, z, A_log, D)]
        y, ys, hl = _scan_ref(*leaves, h0.double())
        (y * dy.double()).sum().backward()
    yg, ysg, hlg = ops.selective_scan(*gl, h0=h0.to(dev()), want_yssm=not pipelined, want_hlast=True, mode=mode)
    if pipelined:
        assert ysg is None
        yg.backward(dy.to(dev(), dtype))
    else:
        torch.autograd.backward([yg, ysg], [dy.to(dev(), dtype), dys.to(dev(), dtype)])
    torch.cuda.synchronize()
    tol = TOL[dtype]
    assert rel_err(yg.float(), y.detach()) < tol, "y"
    if not pipelined:
        assert rel_err(ysg.float(), ys.detach()) < tol, "y_ssm"
    assert rel_err(hlg, hl.detach()) < tol, "h_last"
    names = ["dxa", "ddlog", "dBC", "dz", "dA_log", "dD"]
    for n, a, b in zip(names, gl, leaves):
        assert rel_err(a.grad.float(), b.grad) < (tol if n not in ("dA_log", "dD", "ddlog") else max(tol, 2e-3 if dtype == torch.bfloat16 else tol)), n


def test_selective_scan_modes_agree_and_deterministic():
    from apertis_llm_b200 import ops
    B, L, H = 2, 3000, 8
    Di = 16 * H
    g = torch.Generator().manual_seed(5)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev(), torch.bfloat16)
    xa, z, BC, dlog = mk(B, L, Di), mk(B, L, Di), mk(B, L, 2 * Di), (torch.randn(B, L, H, generator=g) - 3).to(dev(), torch.bfloat16)
    A_log = (torch.rand(H, 16, generator=g) * 0.6 - 0.7).to(dev())
    D = torch.ones(Di, device=dev())
    # each schedule composes the tile aggregates in a fixed order: bitwise reproducible; the two schedules associate the
    # cross-tile products differently (the two-pass combine works on segments), so they agree to fp32 rounding only
    outs = [ops.selective_scan(xa, dlog, BC, z, A_log, D, mode=m)[0] for m in (1, 1, 0, 0, 2, 2)]
    assert torch.equal(outs[0], outs[1]), "two-pass scan is not bitwise deterministic"
    assert torch.equal(outs[2], outs[3]), "single-pass scan is not bitwise deterministic"
    assert torch.equal(outs[4], outs[5]), "pipelined scan is not bitwise deterministic"
    assert rel_err(outs[0].float(), outs[2].float()) < 8e-3, "single-pass and two-pass scans differ"   # one bf16 ulp
    assert rel_err(outs[0].float(), outs[4].float()) < 8e-3, "pipelined and two-pass scans differ"
    xa32, z32, BC32, dl32 = xa.float(), z.float(), BC.float(), dlog.float()
    o32 = [ops.selective_scan(xa32, dl32, BC32, z32, A_log, D, mode=m)[0] for m in (1, 1, 0, 0, 2, 2)]
    assert torch.equal(o32[0], o32[1]) and torch.equal(o32[2], o32[3]) and torch.equal(o32[4], o32[5])
    assert rel_err(o32[0], o32[2]) < 1e-5 and rel_err(o32[0], o32[4]) < 1e-5


@pytest.mark.parametrize("B,L,H", [(1, 50000, 32), (2, 9000, 16), (8, 4096, 12)])
def test_selective_scan_pipelined_long(B, L, H):
    """Steady state of the pipelined schedule (many super-tiles per CTA, ticket recycling, repeated launches on one
    workspace) against the two-pass schedule: outputs and every gradient, twice in a row, bitwise repeatable."""
    from apertis_llm_b200 import _lib, ops
    Di = 16 * H
    g = torch.Generator().manual_seed(L)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev(), torch.bfloat16)
    xa, z, BC = mk(B, L, Di), mk(B, L, Di), mk(B, L, 2 * Di) * 0.5
    dlog = (torch.randn(B, L, H, generator=g) - 3).to(dev(), torch.bfloat16)
    A_log = (torch.rand(H, 16, generator=g) * 0.6 - 0.7).to(dev())
    D = (1.0 + 0.1 * torch.randn(Di, generator=g)).to(dev())
    dy = mk(B, L, Di)
    res = {}
    for mode in (_lib.SCAN_TWO_PASS, _lib.SCAN_PIPELINED, _lib.SCAN_PIPELINED):
        leaves = [t.detach().clone().requires_grad_(True) for t in (xa, dlog, BC, z, A_log, D)]
        y, _, hl = ops.selective_scan(*leaves, want_hlast=True, mode=mode)
        y.backward(dy)
        torch.cuda.synchronize()
        res.setdefault(mode, []).append([y.detach().float(), hl] + [t.grad.float() for t in leaves])
    ref, (p1, p2) = res[_lib.SCAN_TWO_PASS][0], res[_lib.SCAN_PIPELINED]
    names = ["y", "h_last", "dxa", "ddlog", "dBC", "dz", "dA_log", "dD"]
    for n, a, b, c in zip(names, ref, p1, p2):
        assert torch.equal(b, c), f"{n}: pipelined scan is not bitwise repeatable"
        assert rel_err(b, a) < (2e-3 if n in ("dA_log", "dD", "h_last") else 1.6e-2), n


def test_selective_scan_pipelined_cuda_graph():
    """The pipelined schedule keeps its launch epoch in the workspace: a captured forward + backward replays correctly."""
    from apertis_llm_b200 import _lib, ops
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
        y = ops.selective_scan(*leaves, mode=_lib.SCAN_PIPELINED)[0]
        grads = torch.autograd.grad(y, leaves, dy)
        return [y] + list(grads)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                                # eager: allocates this stream's workspace before the capture
        ref = [t.clone() for t in step()]
        torch.cuda.synchronize()
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
            assert torch.equal(a, b), "graph replay of the pipelined scan differs from the eager launch"


# ------------------------------------------------------------------------------------------------
# router, top-k, plan
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,Dm,E,K", [(96, 64, 8, 2), (513, 704, 8, 2), (200, 96, 4, 2), (64, 256, 10, 3)])
def test_router_fwd(S, Dm, E, K):
    from apertis_llm_b200 import ops
    sd = O.make_layer_params(Dm, max(1, Dm // 64), 128, E, seed=S)
    _, moe, _ = O.split_layer_params(sd)
    x, noise = O.make_inputs(1, S, Dm, E, seed=S)
    x2 = x.reshape(S, Dm)
    logits, gates, probs, idx, w = O.moe_router(moe, x2, eps=1e-12, noise=noise, alpha=0.1, K=K)
    ns = F.softplus(moe["w_noise"]) * 0.1
    d = dev()
    r = ops.moe_route(x2.to(d), moe["router_norm.weight"].to(d), moe["router_norm.bias"].to(d), 1e-12, moe["router.weight"].to(d),
                      moe["router.bias"].to(d), noise.to(d), ns.to(d), K)
    assert rel_err(r["logits"], logits) < 1e-5
    assert rel_err(r["gates"], gates) < 1e-5
    assert np.array_equal(r["idx"].cpu().numpy(), idx.numpy().astype(np.int32)), "router indices differ from the oracle"
    assert rel_err(r["w"], w) < 1e-5
    assert rel_err(r["lse"], torch.logsumexp(logits, -1)) < 1e-5
    xn_mean = x2.mean(-1)
    assert rel_err(r["stats"][:, 0], xn_mean) < 1e-4
    aux = r["aux"].cpu()
    assert rel_err(aux[:E], gates.sum(0)) < 1e-5
    cnt = torch.zeros(E).scatter_add_(0, idx.reshape(-1), torch.ones(S * K))
    assert torch.equal(aux[E:2 * E], cnt)


def test_topk_from_logits_bit_exact_with_ties():
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(0)
    S, E, K = 4096, 8, 2
    logits = torch.randn(S, E, generator=g)
    logits[::7] = torch.round(logits[::7])              # many exact ties
    logits[5] = 0.25                                    # an all-equal row
    gates_ref = torch.softmax(logits, -1).numpy()
    gates, idx, probs, w, lse = ops.moe_topk_from_logits(logits.to(dev()), K)
    # selection is defined on the kernel's own gates: ties -> lower expert id
    _, idx_ref = O.topk_lowest_index(gates.cpu().numpy(), K)
    assert np.array_equal(idx.cpu().numpy(), idx_ref.astype(np.int32))
    assert rel_err(gates, gates_ref) < 1e-6
    assert idx[5].tolist() == [0, 1]


@pytest.mark.parametrize("S,E,K,cap,ties,inactive", [(96, 8, 2, 15, False, None), (4096, 8, 2, 640, False, None),
                                                      (1000, 4, 2, 300, True, None), (777, 10, 3, 200, False, [3]),
                                                      (512, 8, 2, 512, False, None), (300, 8, 2, 1, True, [0, 5]),
                                                      (60000, 2, 2, 20000, False, None),      # candidates overflow the smem list
                                                      (32768, 8, 2, 5120, True, None)])
def test_plan_bit_exact(S, E, K, cap, ties, inactive):
    from apertis_llm_b200 import ops
    rng = np.random.default_rng(S + cap)
    idx = np.stack([rng.permutation(E)[:K] for _ in range(S)]).astype(np.int32)
    idx[: S // 3, 0] = 1 % E                            # skew: force overflow on one expert
    w = rng.random((S, K)).astype(np.float32) * 0.9 + 0.05
    if ties:
        w = np.round(w * 8) / 8 + 0.0625
    active = None
    if inactive:
        active = np.ones(E, dtype=bool)
        active[inactive] = False
    kept, counts, groups = O.moe_plan(idx, w, E, cap, active)
    d = dev()
    act_t = torch.from_numpy(active.astype(np.int32)).to(d) if active is not None else None
    p = ops.moe_plan(torch.from_numpy(idx).to(d), torch.from_numpy(w).to(d), E, cap, act_t)
    torch.cuda.synchronize()
    assert np.array_equal(p["counts"].cpu().numpy(), counts.astype(np.int32)), "per-expert counts"
    row_of = p["row_of"].cpu().numpy()
    assert np.array_equal(row_of >= 0, kept), "kept (token, slot) sets"
    seg = p["seg_off"].cpu().numpy()
    RA = _lib_row_align()
    assert seg[0] == 0 and np.all(np.diff(seg) == (counts + RA - 1) // RA * RA)
    n_rows = p["n_rows"].cpu().numpy()
    assert n_rows[0] == seg[-1] and n_rows[1] == counts.sum()
    tok, slot = p["tok_of_row"].cpu().numpy(), p["slot_of_row"].cpu().numpy()
    te = p["tile_expert"].cpu().numpy()
    # permutation is a bijection between kept pairs and non-padding rows, rows live in their expert's segment
    rows = row_of[kept]
    assert len(np.unique(rows)) == rows.size
    s_idx, k_idx = np.nonzero(kept)
    assert np.array_equal(tok[rows], s_idx) and np.array_equal(slot[rows], k_idx)
    e_of = idx[s_idx, k_idx]
    assert np.all(rows >= seg[e_of]) and np.all(rows < seg[e_of] + counts[e_of])
    assert (tok[: n_rows[0]] >= 0).sum() == counts.sum()
    for t in range(n_rows[0] // RA):
        assert seg[te[t]] <= t * RA < seg[te[t] + 1]
    assert np.all(te[n_rows[0] // RA:] == -1)


# ------------------------------------------------------------------------------------------------
# grouped GEMM (tcgen05)
# ------------------------------------------------------------------------------------------------
def _lib_row_align():
    from apertis_llm_b200 import _lib
    assert _lib.query("ab_gemm_row_tile") == _lib.ROW_ALIGN
    return _lib.ROW_ALIGN


def _fake_plan(counts, d):
    E = len(counts)
    RA = _lib_row_align()
    pad = [(c + RA - 1) // RA * RA for c in counts]
    seg = np.concatenate([[0], np.cumsum(pad)]).astype(np.int32)
    max_rows = int(seg[-1]) + 2 * RA                     # spare tiles past the end must be skipped
    te = -np.ones(max_rows // RA, dtype=np.int32)
    for e in range(E):
        te[seg[e] // RA: seg[e + 1] // RA] = e
    valid = np.zeros(max_rows, dtype=bool)
    for e in range(E):
        valid[seg[e]: seg[e] + counts[e]] = True
    plan = dict(tile_expert=torch.from_numpy(te).to(d), n_rows=torch.tensor([int(seg[-1]), int(sum(counts))], dtype=torch.int32, device=d),
                seg_off=torch.from_numpy(seg).to(d), max_rows=max_rows)
    return plan, seg, te, valid


@pytest.mark.parametrize("N,K,counts", [(128, 64, [128, 0, 200]), (2816, 704, [640, 1, 300, 0, 129, 640, 77, 5]),
                                        (704, 2816, [300, 640]), (160, 96, [50, 70, 0, 128]), (256, 1600, [130, 10])])
@pytest.mark.parametrize("mode", ["nt", "nn"])
def test_grouped_gemm_rows(mode, N, K, counts):
    from apertis_llm_b200 import _lib, ops
    d = dev()
    E = len(counts)
    plan, seg, te, valid = _fake_plan(counts, d)
    g = torch.Generator().manual_seed(N + K)
    A = (torch.randn(plan["max_rows"], K, generator=g) * 0.5).to(torch.bfloat16)
    A[~torch.from_numpy(valid)] = 0
    W = (torch.randn(E, N, K, generator=g) * 0.05).to(torch.bfloat16) if mode == "nt" else (torch.randn(E, K, N, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(E, N, generator=g) * 0.1
    Ad, Wd, bd = A.to(d), W.to(d), bias.to(d)
    total = int(seg[-1])
    ref = torch.zeros(total, N, dtype=torch.float32)
    for e in range(E):
        a = A[seg[e]:seg[e + 1]].float()
        ref[seg[e]:seg[e + 1]] = a @ (W[e].float().t() if mode == "nt" else W[e].float())
    # plain
    c = ops.grouped_gemm(mode, Ad, Wd, plan, N, K, E, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(c[:total], ref) < 2e-5, "fp32 accumulate of bf16 operands"
    # bias + gelu with bf16 outputs
    RA = _lib_row_align()
    erow = torch.from_numpy(np.repeat(te[: total // RA], RA)).long()
    pre = (ref + bias[erow]).to(torch.bfloat16).float()
    h, hpre = ops.grouped_gemm(mode, Ad, Wd, plan, N, K, E, bias=bd, epi=_lib.EPI_BIAS_ACT, act=0, out_dtype=torch.bfloat16, want_c2=True)
    assert rel_err(hpre[:total].float(), pre) < 1e-2
    assert rel_err(h[:total].float(), F.gelu(pre)) < 1e-2
    # d-activation epilogue against a saved pre-activation
    aux = torch.randn(plan["max_rows"], N, generator=g)      # saved pre-activation has the output dtype
    dact = ops.grouped_gemm(mode, Ad, Wd, plan, N, K, E, aux=aux.to(d), epi=_lib.EPI_DACT, act=0, out_dtype=torch.float32)
    xg = aux[:total].float().requires_grad_(True)
    F.gelu(xg).sum().backward()
    assert rel_err(dact[:total], ref * xg.grad) < 1e-4


@pytest.mark.parametrize("M,N,counts", [(128, 64, [128, 0, 200]), (2816, 704, [640, 1, 300]), (704, 2816, [300, 640]),
                                        (160, 96, [50, 70, 0, 128])])
def test_grouped_gemm_tn(M, N, counts):
    from apertis_llm_b200 import ops
    d = dev()
    E = len(counts)
    plan, seg, te, valid = _fake_plan(counts, d)
    g = torch.Generator().manual_seed(M + N)
    A = (torch.randn(plan["max_rows"], M, generator=g) * 0.5).to(torch.bfloat16)
    Bm = (torch.randn(plan["max_rows"], N, generator=g) * 0.5).to(torch.bfloat16)
    A[~torch.from_numpy(valid)] = 0                       # padding rows of one operand are zero by construction
    out = ops.grouped_gemm_tn(A.to(d), Bm.to(d), plan["seg_off"], M, N, E)
    torch.cuda.synchronize()
    for e in range(E):
        ref = A[seg[e]:seg[e + 1]].float().t() @ Bm[seg[e]:seg[e + 1]].float()
        assert rel_err(out[e], ref) < 2e-5 if counts[e] else float(out[e].abs().max()) == 0.0, e


@pytest.mark.parametrize("W,El,M,N", [(8, 1, 704, 2816), (8, 1, 2816, 704), (4, 2, 704, 1408), (2, 4, 160, 96)])
def test_grouped_gemm_tn_source_blocks(W, El, M, N):
    """Expert-parallel receive layout: expert e's rows are W blocks (one per source rank) of `seg` rows each, block s at
    s * El * seg + e * seg.  With one or two local experts the output has fewer tiles than CTA pairs and the library cuts
    the contraction over the source blocks (ab_grouped_gemm_tn_workspace_bytes > 0): same sums, fixed order."""
    from apertis_llm_b200 import _lib, ops
    d = dev()
    RA = _lib_row_align()
    seg = 2 * RA
    rows = W * El * seg
    g = torch.Generator().manual_seed(W * 100 + El)
    A = (torch.randn(rows, M, generator=g) * 0.5).to(torch.bfloat16)
    Bm = (torch.randn(rows, N, generator=g) * 0.5).to(torch.bfloat16)
    valid = torch.zeros(rows, dtype=torch.bool)
    cnt = torch.randint(1, seg + 1, (W, El), generator=g)
    for s_ in range(W):
        for e in range(El):
            valid[(s_ * El + e) * seg: (s_ * El + e) * seg + int(cnt[s_, e])] = True
    A[~valid] = 0
    seg_off = (torch.arange(El + 1, dtype=torch.int32) * seg).to(d)
    if W == 8 and El == 1:
        assert _lib.query("ab_grouped_gemm_tn_workspace_bytes", M, N, El, W) > 0, "one expert per rank: cut over the source blocks"
    out = ops.grouped_gemm_tn(A.to(d), Bm.to(d), seg_off, M, N, El, nsrc=W, src_stride=El * seg)
    out2 = ops.grouped_gemm_tn(A.to(d), Bm.to(d), seg_off, M, N, El, nsrc=W, src_stride=El * seg)
    torch.cuda.synchronize()
    assert torch.equal(out, out2), "bitwise repeatable"
    Av, Bv = A.float().view(W, El, seg, M), Bm.float().view(W, El, seg, N)
    for e in range(El):
        ref = torch.einsum("srm,srn->mn", Av[:, e], Bv[:, e])
        assert rel_err(out[e], ref) < 2e-5, e


def test_ep_peer_kernels_two_ranks_emulated_on_one_gpu():
    """The peer-memory kernels of expert parallelism (ab_ep_permute_ln, ab_ep_grouped_gemm_nt / nn, ab_ep_unpermute_bwd)
    with two ranks emulated on ONE device: each 'rank' gets its own receive buffers and the address table lists both.  A
    rank's producers must leave exactly the rows, at exactly the places, that the local kernels plus an all-to-all would
    (the multi-GPU NCCL / NVLink runs of the same kernels are tests/test_ep.py and bench.py's ep_parity)."""
    import ctypes
    from apertis_llm_b200 import _lib, ops
    from apertis_llm_b200._lib import call, dt, ptr, stream_ptr
    d = dev()
    RA = _lib_row_align()
    W, El, E, K, S, Dm, I = 2, 2, 4, 2, 600, 96, 160
    cap = 260
    seg = (cap + RA - 1) // RA * RA
    rpp = El * seg                       # rows per peer
    rows = W * rpp
    g = torch.Generator().manual_seed(11)
    table = lambda bufs: (ctypes.c_uint64 * W)(*[b.data_ptr() for b in bufs])
    xn_recv = [torch.full((rows, Dm), float("nan"), dtype=torch.bfloat16, device=d) for _ in range(W)]
    y_in = [torch.full((rows, Dm), float("nan"), dtype=torch.bfloat16, device=d) for _ in range(W)]
    dy_recv = [torch.full((rows, Dm), float("nan"), dtype=torch.bfloat16, device=d) for _ in range(W)]
    local = []
    for rank in range(W):
        x2 = torch.randn(S, Dm, generator=g).to(d)
        idx = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(S)]).to(torch.int32).to(d)
        w = (torch.rand(S, K, generator=g) * 0.9 + 0.05).to(d)
        stats = torch.stack([x2.mean(-1), (x2.var(-1, unbiased=False) + 1e-12).rsqrt()], -1).contiguous()
        ln_w, ln_b = (1 + 0.1 * torch.randn(E, Dm, generator=g)).to(d), (0.1 * torch.randn(E, Dm, generator=g)).to(d)
        plan = ops.moe_plan(idx, w, E, cap, None, fixed_seg=seg)
        # local kernel, then what an all-to-all would deliver: block `dst` of the local layout -> block `rank` of rank dst
        xn = torch.empty(rows, Dm, dtype=torch.bfloat16, device=d)
        call("ab_moe_permute_ln", ptr(x2), ptr(stats), ptr(ln_w), ptr(ln_b), ptr(plan["tok_of_row"]), ptr(plan["tile_expert"]),
             ptr(plan["n_rows"]), ptr(xn), Dm, RA, rows, dt(x2), dt(xn), stream_ptr())
        call("ab_ep_permute_ln", ptr(x2), ptr(stats), ptr(ln_w), ptr(ln_b), ptr(plan["tok_of_row"]), ptr(plan["tile_expert"]),
             ptr(plan["n_rows"]), table(xn_recv), W, rank, rpp, Dm, RA, rows, dt(x2), dt(xn), stream_ptr())
        dout = torch.randn(S, Dm, generator=g).to(d)
        yloc = (torch.randn(rows, Dm, generator=g) * 0.5).to(torch.bfloat16).to(d)
        dy = torch.empty(rows, Dm, dtype=torch.bfloat16, device=d)
        dwr, dwr2 = torch.empty(rows, device=d), torch.empty(rows, device=d)
        call("ab_moe_unpermute_bwd", ptr(dout), ptr(yloc), ptr(w), ptr(plan["tok_of_row"]), ptr(plan["slot_of_row"]), ptr(plan["n_rows"]),
             ptr(dy), ptr(dwr), 0.0, None, K, Dm, rows, dt(dout), dt(yloc), dt(dy), stream_ptr())
        call("ab_ep_unpermute_bwd", ptr(dout), ptr(yloc), ptr(w), ptr(plan["tok_of_row"]), ptr(plan["slot_of_row"]), ptr(plan["n_rows"]),
             table(dy_recv), W, rank, rpp, ptr(dwr2), 0.0, None, K, Dm, rows, dt(dout), dt(yloc), dt(dy), stream_ptr())
        assert torch.equal(dwr, dwr2)
        local.append((xn, dy))
    torch.cuda.synchronize()
    for owner in range(W):
        for src in range(W):
            blk = slice(src * rpp, (src + 1) * rpp)
            sent = slice(owner * rpp, (owner + 1) * rpp)
            assert torch.equal(xn_recv[owner][blk], local[src][0][sent]), ("dispatch", owner, src)
            assert torch.equal(dy_recv[owner][blk], local[src][1][sent]), ("dY dispatch", owner, src)
    # the owners' GEMMs return their result rows into the source ranks' buffers
    tile_expert = torch.arange(El, dtype=torch.int32, device=d).repeat_interleave(seg // RA).repeat(W)
    n_rows = torch.full((2,), rows, dtype=torch.int32, device=d)
    rplan = dict(tile_expert=tile_expert, n_rows=n_rows)
    for owner in range(W):
        A = xn_recv[owner]
        Wt = (torch.randn(El, Dm, Dm, generator=g) * 0.1).to(torch.bfloat16).to(d)
        bias = torch.randn(El, Dm, generator=g).to(d)
        ref = ops.grouped_gemm("nt", A, Wt, rplan, Dm, Dm, El, bias=bias, epi=_lib.EPI_BIAS)
        call("ab_ep_grouped_gemm_nt", ptr(A), ptr(Wt), ptr(bias), None, table(y_in), W, owner, rpp, ptr(tile_expert), ptr(n_rows), rows,
             Dm, Dm, El, _lib.EPI_BIAS, 0, dt(torch.bfloat16), stream_ptr())
        ref2 = ops.grouped_gemm("nn", A, Wt, rplan, Dm, Dm, El)
        torch.cuda.synchronize()
        for src in range(W):
            assert torch.equal(y_in[src][owner * rpp:(owner + 1) * rpp], ref[src * rpp:(src + 1) * rpp]), ("combine", owner, src)
        call("ab_ep_grouped_gemm_nn", ptr(A), ptr(Wt), None, None, table(y_in), W, owner, rpp, ptr(tile_expert), ptr(n_rows), rows,
             Dm, Dm, El, _lib.EPI_NONE, 0, dt(torch.bfloat16), stream_ptr())
        torch.cuda.synchronize()
        for src in range(W):
            assert torch.equal(y_in[src][owner * rpp:(owner + 1) * rpp], ref2[src * rpp:(src + 1) * rpp]), ("dXn return", owner, src)


# ------------------------------------------------------------------------------------------------
# block-wrapper LayerNorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,Dm", [(96, 64), (1000, 704), (257, 96)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_layer_norm(S, Dm, out_dtype):
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(S)
    x = torch.randn(S, Dm, generator=g) * 2 + 0.5
    w, b = 1 + 0.1 * torch.randn(Dm, generator=g), 0.1 * torch.randn(Dm, generator=g)
    dy = torch.randn(S, Dm, generator=g).to(out_dtype).float()
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.layer_norm(xr, (Dm,), wr, br, 1e-12)
    ref.backward(dy)
    xg, wg, bg = (t.to(dev()).requires_grad_(True) for t in (x, w, b))
    out = ops.layer_norm(xg, wg, bg, 1e-12, out_dtype=out_dtype)
    out.backward(dy.to(dev(), out_dtype))
    tol = TOL[out_dtype]
    assert rel_err(out.float(), ref.detach()) < tol
    assert rel_err(xg.grad, xr.grad) < 1e-4 and rel_err(wg.grad, wr.grad) < 1e-4 and rel_err(bg.grad, br.grad) < 1e-4


# ------------------------------------------------------------------------------------------------
# expert-internal dropout fused into the GEMM epilogues (core.py:439)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_grouped_gemm_dropout_mask_consistent(p):
    from apertis_llm_b200 import _lib, ops
    d = dev()
    N, K, counts = 704, 256, [300, 129, 0, 640]
    E = len(counts)
    plan, seg, te, valid = _fake_plan(counts, d)
    g = torch.Generator().manual_seed(7)
    A = (torch.randn(plan["max_rows"], K, generator=g) * 0.5).to(torch.bfloat16).to(d)
    W = (torch.randn(E, N, K, generator=g) * 0.05).to(torch.bfloat16).to(d)
    bias = (torch.randn(E, N, generator=g) * 0.1).to(d)
    seed = torch.tensor([12345, 678], dtype=torch.int32, device=d)
    total = int(seg[-1])
    h0, pre0 = ops.grouped_gemm("nt", A, W, plan, N, K, E, bias=bias, epi=_lib.EPI_BIAS_ACT, act=0, want_c2=True)
    h1, pre1 = ops.grouped_gemm("nt", A, W, plan, N, K, E, bias=bias, epi=_lib.EPI_BIAS_ACT, act=0, want_c2=True, drop_p=p, drop_seed=seed)
    h2, _ = ops.grouped_gemm("nt", A, W, plan, N, K, E, bias=bias, epi=_lib.EPI_BIAS_ACT, act=0, want_c2=True, drop_p=p, drop_seed=seed)
    assert torch.equal(pre0[:total], pre1[:total]) and torch.equal(h1[:total], h2[:total])   # pre-activation untouched; mask reproducible
    h0f, h1f = h0[:total].float(), h1[:total].float()
    keep = h1f != 0
    frac = 1.0 - keep.float().mean().item()
    n = keep.numel()
    assert abs(frac - p) < 5 * (p * (1 - p) / n) ** 0.5 + 2e-3, frac    # gelu(x) == 0 exactly is negligible
    assert rel_err(h1f[keep], (h0f / (1 - p))[keep]) < 1e-2
    # a different seed gives a different mask
    h3, _ = ops.grouped_gemm("nt", A, W, plan, N, K, E, bias=bias, epi=_lib.EPI_BIAS_ACT, act=0, want_c2=True, drop_p=p,
                             drop_seed=torch.tensor([12346, 678], dtype=torch.int32, device=d))
    assert not torch.equal(h1[:total], h3[:total])
    # the dact epilogue regenerates the same mask over the same [rows, N] index space
    dy = (torch.randn(plan["max_rows"], K, generator=g) * 0.5).to(torch.bfloat16).to(d)
    W2 = (torch.randn(E, K, N, generator=g) * 0.05).to(torch.bfloat16).to(d)
    for odt, aux, tol in ((torch.float32, pre1.float(), 1e-5), (torch.bfloat16, pre1, 1e-2)):      # aux has the output's dtype
        g0 = ops.grouped_gemm("nn", dy, W2, plan, N, K, E, aux=aux, epi=_lib.EPI_DACT, act=0, out_dtype=odt).float()
        g1 = ops.grouped_gemm("nn", dy, W2, plan, N, K, E, aux=aux, epi=_lib.EPI_DACT, act=0, out_dtype=odt, drop_p=p, drop_seed=seed).float()
        assert torch.equal((g1[:total] != 0) | (g0[:total] == 0), keep | (g0[:total] == 0))
        assert rel_err(g1[:total][keep], (g0[:total] / (1 - p))[keep]) < tol


def test_grouped_gemm_gelu_accuracy_per_element():
    """The bf16 data path evaluates GELU / GELU' through a tanh-form fit of the normal cdf (csrc/grouped_gemm.cu); the fp32
    path through an erf with |error| < 1.5e-7.  Element by element over pre-activations in [-12, 12]: the bf16 results are
    the exact-erf values to bf16 rounding plus the fit's 2.6e-4 |x|, the fp32 results to 2e-6."""
    from apertis_llm_b200 import _lib, ops
    d = dev()
    RA = _lib_row_align()
    N, K, E = 64, 64, 1
    rows = 4 * RA
    plan = dict(tile_expert=torch.zeros(rows // RA, dtype=torch.int32, device=d), n_rows=torch.tensor([rows, rows], dtype=torch.int32, device=d))
    xs = torch.linspace(-12, 12, rows).to(torch.bfloat16)                 # one pre-activation value per row (bf16-exact)
    A = torch.zeros(rows, K, dtype=torch.bfloat16)
    A[:, 0] = xs
    W = torch.zeros(E, N, K, dtype=torch.bfloat16)
    W[0, :, 0] = 1.0                                                       # pre[r, n] = xs[r]
    bias = torch.zeros(E, N)
    x32 = xs.float()[:, None].expand(rows, N)
    want = F.gelu(x32)
    xg = x32.clone().requires_grad_(True)
    F.gelu(xg).sum().backward()
    want_d = xg.grad
    ones = torch.zeros(rows, K, dtype=torch.bfloat16)
    ones[:, 0] = 1.0                                                       # accumulator of the dact GEMM = 1
    for odt in (torch.bfloat16, torch.float32):
        h, pre = ops.grouped_gemm("nt", A.to(d), W.to(d), plan, N, K, E, bias=bias.to(d), epi=_lib.EPI_BIAS_ACT, act=0, out_dtype=odt, want_c2=True)
        assert torch.equal(pre.float().cpu(), x32)
        dact = ops.grouped_gemm("nt", ones.to(d), W.to(d), plan, N, K, E, aux=x32.to(odt).to(d), epi=_lib.EPI_DACT, act=0, out_dtype=odt)
        torch.cuda.synchronize()
        if odt == torch.float32:
            assert float((h.cpu() - want).abs().max()) < 2e-6 * 12 and float((dact.cpu() - want_d).abs().max()) < 3e-6
        else:
            eh = (h.float().cpu() - want).abs()
            assert bool((eh <= 2.0 ** -8 * want.abs() + 3e-4 * x32.abs() + 1e-4).all()), float(eh.max())
            ed = (dact.float().cpu() - want_d).abs()
            assert bool((ed <= 2.0 ** -8 * want_d.abs() + 1.5e-3).all()), float(ed.max())


@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("sub_dtype", [torch.float32, torch.bfloat16])
def test_dropout_add(p, sub_dtype):
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(3)
    sub = torch.randn(257, 64, generator=g).to(dev(), sub_dtype).requires_grad_(True)
    res = torch.randn(257, 64, generator=g).to(dev()).requires_grad_(True)
    torch.manual_seed(11)
    out = ops.dropout_add(sub, res, p, True)
    dout = torch.randn_like(out)
    out.backward(dout)
    d = (out - res).detach()
    keep = d != 0
    if p == 0:
        assert rel_err(out, sub.float() + res) < (1e-6 if sub_dtype == torch.float32 else 1e-2)
        assert torch.equal(res.grad, dout) and rel_err(sub.grad.float(), dout) < 1e-2
        return
    frac = 1 - keep.float().mean().item()
    assert abs(frac - p) < 0.02
    assert rel_err(d[keep], (sub.detach().float() / (1 - p))[keep]) < 1e-2
    assert torch.equal(res.grad, dout)
    assert rel_err(sub.grad.float()[keep], (dout / (1 - p))[keep]) < 1e-2 and float(sub.grad.float()[~keep].abs().max()) == 0.0
    assert torch.equal(ops.dropout_add(sub, res, p, False), sub.float() + res) or sub_dtype == torch.bfloat16


# ------------------------------------------------------------------------------------------------
# dense GEMMs of the SSM projections (core.py:366-367, 376-383, 397) on the tcgen05 kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("S,N,K", [(4673, 352, 704), (1, 368, 176), (300, 704, 176), (4096, 176, 368), (777, 832, 400), (130, 8, 16)])
def test_dense_gemm_nt_nn_tn(precise, S, N, K):
    """C = A W^T (+bias), dX = dY W (+add), dW = dY^T A against torch fp32 matmuls of the same (bf16-rounded) operands."""
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(S + N + K)
    q = (lambda t: t) if precise else (lambda t: t.to(torch.bfloat16).float())
    A, W, dY = q(torch.randn(S, K, generator=g)), q(torch.randn(N, K, generator=g) * 0.1), q(torch.randn(S, N, generator=g))
    bias, add = torch.randn(N, generator=g), q(torch.randn(S, K, generator=g))
    d = dev()
    tol = 1e-4 if precise else 2e-2
    cdt = torch.float32 if precise else torch.bfloat16
    c = ops.dense_nt(A.to(d, cdt), W.to(d, cdt), precise, bias=bias.to(d))
    assert rel_err(c.float(), A @ W.t() + bias) < tol
    dx = ops.dense_nn(dY.to(d, cdt), W.to(d, cdt), precise, add=add.to(d, cdt))
    assert rel_err(dx.float(), dY @ W + add) < tol
    dw = ops.dense_tn(dY.to(d, cdt), A.to(d, cdt), precise)
    assert dw.dtype == torch.float32 and rel_err(dw, dY.t() @ A) < (1e-4 if precise else 5e-3)


@pytest.mark.parametrize("autocast", [False, True])
def test_linear_autograd_matches_torch(autocast):
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 333, 704, generator=g)
    w1, w2 = torch.randn(176, 704, generator=g) * 0.05, torch.randn(176, 704, generator=g) * 0.05
    dy = torch.randn(2, 333, 352, generator=g)
    xr, w1r, w2r = (t.clone().requires_grad_(True) for t in (x, w1, w2))
    torch.nn.functional.linear(xr, torch.cat([w1r, w2r], 0)).backward(dy)
    xg, w1g, w2g = (t.to(dev()).requires_grad_(True) for t in (x, w1, w2))
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = ops.linear(xg.to(torch.bfloat16) if autocast else xg, [w1g, w2g], precise=not autocast)
    y.backward(dy.to(dev(), y.dtype))
    tol = 2e-2 if autocast else 1e-4
    assert rel_err(xg.grad, xr.grad) < tol and rel_err(w1g.grad, w1r.grad) < tol and rel_err(w2g.grad, w2r.grad) < tol


def test_dt_compose_fwd_bwd():
    from apertis_llm_b200 import ops
    g = torch.Generator().manual_seed(8)
    H, R, Di = 11, 44, 176
    Wp, Wdt = torch.randn(R + 2 * Di, Di, generator=g), torch.randn(H, R, generator=g)
    Wpr, Wdtr = Wp.clone().requires_grad_(True), Wdt.clone().requires_grad_(True)
    ref = torch.cat([Wdtr @ Wpr[:R], torch.zeros(16 - H, Di), Wpr[R:]], 0)
    dW = torch.randn(ref.shape, generator=g)
    ref.backward(dW)
    Wpg, Wdtg = Wp.to(dev()).requires_grad_(True), Wdt.to(dev()).requires_grad_(True)
    out = ops.dt_compose(Wpg, Wdtg, precise=True)
    out.backward(dW.to(dev()))
    assert rel_err(out, ref.detach()) < 1e-5
    assert rel_err(Wpg.grad, Wpr.grad) < 1e-5 and rel_err(Wdtg.grad, Wdtr.grad) < 1e-5
