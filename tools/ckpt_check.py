"""Gradients of one block with and without torch.utils.checkpoint around it, fp32 and bf16 autocast (one GPU, no EP):
how far a different call sequence moves them (python tools/ckpt_check.py)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from torch.utils.checkpoint import checkpoint
from apertis_llm_b200 import ApertisLayerB200, BlockConfig
from oracle import apertis_oracle as O
dev = torch.device("cuda:0")
cfg = BlockConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_experts=8, experts_per_token=2, hidden_dropout_prob=0.0)
sd = O.make_layer_params(128, 2, 256, 8, seed=5)
layer = ApertisLayerB200(cfg); layer.load_state_dict(sd, strict=True); layer = layer.to(dev).train()
x, noise = O.make_inputs(2, 160, 128, 8, seed=40)
layer.feed_forward.ffn._draw_noise = lambda S_, E_, device: noise.to(device)
def run(ck, ac):
    for p in layer.parameters(): p.grad = None
    xg = x.to(dev).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
        if ck: out, lb, rz = checkpoint(lambda t: tuple(layer(t)[i] for i in (0, 3, 4)), xg, use_reentrant=False)
        else:
            o = layer(xg); out, lb, rz = o[0], o[3], o[4]
    O.block_loss(out, lb, rz).backward()
    return {n: p.grad.clone() for n, p in layer.named_parameters()}, xg.grad.clone()
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
for ac in (False, True):
    g0, d0 = run(False, ac); g0b, d0b = run(False, ac); g1, d1 = run(True, ac)
    worst = sorted(((rel(g1[n], g0[n]), n) for n in g0), reverse=True)[:4]
    rr = sorted(((rel(g0b[n], g0[n]), n) for n in g0), reverse=True)[:2]
    print("autocast", ac, "plain vs plain", rr, "| ckpt vs plain dx", rel(d1, d0), worst)
