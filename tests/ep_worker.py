"""Multi-process worker for the expert-parallel tests (launched by torchrun or mp.spawn).

  --mode gpu : NCCL, one rank per GPU: EP layer vs the same layer with all experts local, on each rank's own tokens
  --mode cpu : gloo on CPU: exchange primitive, layout maps and a simulation of the EP data flow with the oracle's
               expert math (no kernels), state_dict sharding hooks
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import apertis_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def run_cpu(rank, world):
    import torch.nn.functional as F
    from apertis_llm_b200 import AdaptiveExpertSystem, BlockConfig, ep
    g = dist.group.WORLD
    # 1. exchange primitive
    send = torch.arange(world * 3 * 2, dtype=torch.float32).view(world, 3, 2) + 100 * rank
    recv = ep.all_to_all_equal(send.contiguous(), g)
    for s in range(world):
        exp = torch.arange(world * 3 * 2, dtype=torch.float32).view(world, 3, 2)[rank] + 100 * s
        assert torch.equal(recv[s], exp), (rank, s)
    # 2. EP data flow simulated with the oracle's math
    Dm, H, I, E, K, B, L = 32, 2, 64, 4, 2, 1, 40
    El = E // world
    sd = O.make_layer_params(Dm, H, I, E, seed=9)
    _, moe, _ = O.split_layer_params(sd)
    x, noise = O.make_inputs(B, L, Dm, E, seed=20 + rank)
    S = B * L
    x2 = x.reshape(S, Dm)
    ref, lb, rz, parts = O.moe_forward(moe, x, E=E, K=K, training=True, noise=noise, return_parts=True)
    cap = parts["cap"]
    seg = ep.segment_rows(cap)
    kept, counts, groups = parts["kept"], parts["counts"], None
    idx, w = parts["idx"].numpy(), parts["w"].detach().numpy()
    # fixed-segment permuted layout: expert e rows at [e*seg, e*seg+count_e), token order inside (slot-major)
    rows = torch.zeros(E * seg, Dm)
    row_of = -np.ones((S, K), dtype=np.int64)
    fill = np.zeros(E, dtype=np.int64)
    for k in range(K):
        for s in range(S):
            if kept[s, k]:
                e = idx[s, k]
                r = e * seg + fill[e]
                fill[e] += 1
                row_of[s, k] = r
                xe = x2[s]
                rows[r] = F.layer_norm(xe, (Dm,), moe[f"experts.{e}.0.weight"], moe[f"experts.{e}.0.bias"], 1e-12)
    assert np.array_equal(fill, counts)
    recvd = ep.all_to_all_equal(rows.view(world, El * seg, Dm).contiguous(), g).view(E * seg, Dm)
    te = ep.recv_tile_expert(world, El, seg)
    yr = torch.zeros_like(recvd)
    RA = ep.ROW_ALIGN
    for t in range(recvd.shape[0] // RA):
        e = rank * El + int(te[t])
        blk = recvd[t * RA:(t + 1) * RA]
        hdn = F.gelu(F.linear(blk, moe[f"experts.{e}.1.weight"], moe[f"experts.{e}.1.bias"]))
        yr[t * RA:(t + 1) * RA] = F.linear(hdn, moe[f"experts.{e}.4.weight"], moe[f"experts.{e}.4.bias"])
    y = ep.all_to_all_equal(yr.view(world, El * seg, Dm).contiguous(), g).view(E * seg, Dm)
    out = torch.zeros(S, Dm)
    for k in range(K):
        for s in range(S):
            if row_of[s, k] >= 0:
                out[s] += y[row_of[s, k]] * w[s, k]
    assert rel(out, ref.reshape(S, Dm)) < 1e-5, rel(out, ref.reshape(S, Dm))
    so = ep.local_seg_off(El, seg)
    assert so.tolist() == [i * seg for i in range(El + 1)]
    # 3. state_dict sharding: a rank keeps only its experts, under the reference's global key names
    m = AdaptiveExpertSystem(BlockConfig(hidden_size=Dm, num_attention_heads=H, intermediate_size=I, num_experts=E), ep_group=g)
    m.load_state_dict(moe, strict=True)
    assert m.expert_w1.shape == (El, I, Dm)
    for le in range(El):
        assert torch.equal(m.expert_w1[le], moe[f"experts.{rank * El + le}.1.weight"])
    keys = set(m.state_dict())
    assert f"experts.{rank * El}.1.weight" in keys and f"experts.{(rank * El + El) % E}.1.weight" not in keys or world == 1
    assert len(ep.replicated_parameters(m)) == 5      # w_noise, router_norm.{w,b}, router.{w,b}
    print(f"rank {rank}: cpu ep host logic ok (cap {cap}, seg {seg}, counts {counts.tolist()})")


def run_gpu(rank, world, args):
    from apertis_llm_b200 import ApertisLayerB200, BlockConfig, ep
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    g = dist.group.WORLD
    Dm, H, I, E, K, B, L = args.dm, args.heads, args.inter, 8, 2, 2, args.seq
    sd = O.make_layer_params(Dm, H, I, E, seed=5)
    cfg = BlockConfig(hidden_size=Dm, num_attention_heads=H, intermediate_size=I, num_experts=E, experts_per_token=K,
                      hidden_dropout_prob=0.0)
    El = E // world
    worst = 0.0
    for autocast in (False, True):
        x, noise = O.make_inputs(B, L, Dm, E, seed=40 + rank)
        res = []
        for use_ep in (False, True):
            layer = ApertisLayerB200(cfg, ep_group=g if use_ep else None)
            layer.load_state_dict(sd, strict=True)
            layer = layer.to(dev).train()
            layer.feed_forward.ffn._draw_noise = lambda S_, E_, device: noise.to(device)
            xg = x.to(dev).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                out, _, _, lb, rz = layer(xg)
            O.block_loss(out, lb, rz).backward()
            res.append((layer, out.detach(), xg.grad.detach()))
        (lr, out_r, dx_r), (le, out_e, dx_e) = res
        tol = 2e-2 if autocast else 1e-4
        errs = {"out": rel(out_e, out_r), "dx": rel(dx_e, dx_r)}
        fr, fe = lr.feed_forward.ffn, le.feed_forward.ffn
        assert torch.equal(fr.last_counts, fe.last_counts)
        for name in fe._STACKED:
            gr = getattr(fr, name).grad.clone()
            dist.all_reduce(gr)                                 # sum over ranks of the replicated-expert gradients
            gr = gr[rank * El:(rank + 1) * El] / world          # DDP mean, this rank's experts
            errs[name] = rel(getattr(fe, name).grad, gr)
        for (n1, p1), (n2, p2) in zip(lr.named_parameters(), le.named_parameters()):
            if ".expert_" in n1:
                continue
            errs[n1] = rel(p2.grad, p1.grad)
        bad = {k: v for k, v in errs.items() if not v < tol}
        assert not bad, (rank, autocast, bad)
        worst = max(worst, max(errs.values()))
        # activation checkpointing around the EP layer (the reference trainer's default, core.py:1258-1272): the
        # recomputation rewrites the peer-mapped buffers before the backward reads them; gradients must not move beyond
        # the run-to-run reduction-order noise of a different call sequence (measured on one GPU without EP,
        # tools/ckpt_check.py: 8e-6 in fp32, 6e-3 under bf16 on the SSM parameters) - a stale buffer would be O(1)
        from torch.utils.checkpoint import checkpoint
        g_plain = {n: p.grad.clone() for n, p in le.named_parameters()}
        for p in le.parameters():
            p.grad = None
        xc = x.to(dev).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out_c, lb_c, rz_c = checkpoint(lambda t: tuple(le(t)[i] for i in (0, 3, 4)), xc, use_reentrant=False)
        O.block_loss(out_c, lb_c, rz_c).backward()
        assert rel(xc.grad, dx_e) < tol, (rank, autocast, "checkpoint dx", rel(xc.grad, dx_e))
        for n, p in le.named_parameters():
            assert rel(p.grad, g_plain[n]) < tol, (rank, autocast, "checkpoint", n, rel(p.grad, g_plain[n]))
        # evaluation: no capacity limit, the exchange segments are sized from the counts reduced over the ranks
        lr.eval(); le.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            ev_r, ev_e = lr(x.to(dev))[0], le(x.to(dev))[0]
        assert rel(ev_e, ev_r) < tol, (rank, autocast, "eval", rel(ev_e, ev_r))
        assert torch.equal(lr.feed_forward.ffn.last_counts, le.feed_forward.ffn.last_counts)
    transport = os.environ.get("APERTIS_B200_EP", "auto")
    if transport == "peer":
        assert ep._peer_cache, "the peer-memory transport was requested but never used"
    if transport == "nccl":
        assert not ep._peer_cache
    print(f"rank {rank}: gpu ep ok, world {world}, transport {transport} ({len(ep._peer_cache)} peer buffer sets), worst rel err {worst:.2e}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cpu")
    ap.add_argument("--dm", type=int, default=128)
    ap.add_argument("--heads", type=int, default=2)
    ap.add_argument("--inter", type=int, default=256)
    ap.add_argument("--seq", type=int, default=160)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if args.mode == "gpu":
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=dev)
        run_gpu(rank, world, args)
    else:
        dist.init_process_group("gloo")
        run_cpu(rank, world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
