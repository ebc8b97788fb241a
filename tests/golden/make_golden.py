#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
modules from /root/reference (build container only; the GPU box has no reference).

    python tests/golden/make_golden.py

Weights and inputs come from ``oracle.apertis_oracle.make_layer_params / make_inputs``
(seeded CPU generators) and are loaded into the reference's ``ApertisLayer`` with
``load_state_dict(strict=True)``; the routing noise the reference draws with
``torch.randn_like`` (core.py:487) and the expert-dropout permutation (core.py:519) are
supplied by patching those two torch functions for the duration of the call, so that the
fixture records exactly which random numbers were consumed.  Nothing in the reference is
edited or copied.

Each fixture is an .npz holding the case spec (json), the reference outputs, losses,
input gradient, parameter gradients (full for small cases, a strided sample + sums for
the C1-dims case) and the routing artefacts (top-k indices, normalised weights).
"""
import contextlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import apertis_oracle as O  # noqa: E402
from src.model.core import ApertisConfig, ApertisLayer  # noqa: E402  (the reference)

torch.set_num_threads(4)

SAMPLE = 2048


def sample(t: torch.Tensor) -> np.ndarray:
    """Strided sample of a big tensor (first element of every stride-th)."""
    f = t.detach().reshape(-1)
    if f.numel() <= SAMPLE:
        return f.numpy().copy()
    step = f.numel() // SAMPLE
    return f[::step][:SAMPLE].numpy().copy()


@contextlib.contextmanager
def patched_rng(noise, perm):
    orig_rl, orig_rp = torch.randn_like, torch.randperm

    def fake_randn_like(t, *a, **k):
        assert noise is not None and tuple(t.shape) == tuple(noise.shape), (t.shape,)
        return noise.to(t.dtype)

    def fake_randperm(n, *a, **k):
        assert perm is not None and n == len(perm)
        return torch.tensor(perm, dtype=torch.long)

    torch.randn_like, torch.randperm = fake_randn_like, fake_randperm
    try:
        yield
    finally:
        torch.randn_like, torch.randperm = orig_rl, orig_rp


def ref_layer(spec):
    cfg = ApertisConfig(hidden_size=spec["Dm"], num_attention_heads=spec["H"], intermediate_size=spec["I"],
                        num_hidden_layers=1, attention_type="selective_ssm", use_expert_system=True,
                        num_experts=spec["E"], experts_per_token=spec["K"], vocab_size=64,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                        hidden_act=spec.get("act", "gelu"))
    layer = ApertisLayer(cfg)
    sd = O.make_layer_params(spec["Dm"], spec["H"], spec["I"], spec["E"], seed=spec["seed"])
    # the reference layer also owns a RotaryEmbedding buffer-less module; strict load of our keys
    missing, unexpected = layer.load_state_dict(sd, strict=False)
    assert not unexpected and all("rope" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    return layer, sd


def case_block(name, spec, full_grads=True):
    layer, sd = ref_layer(spec)
    x, noise = O.make_inputs(spec["B"], spec["L"], spec["Dm"], spec["E"], seed=spec["seed"])
    training = spec.get("training", True)
    layer.train(training)
    x = x.requires_grad_(True)
    perm = spec.get("perm")
    with patched_rng(noise if training else None, perm):
        out, _, _, lb, rz = layer(x)
    arrs = {"out": out.detach().numpy(), "lb": lb.detach().numpy(), "rz": rz.detach().numpy()}
    if training:
        O.block_loss(out, lb, rz).backward()
        arrs["dx"] = x.grad.numpy()
        for k, p in layer.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if full_grads:
                arrs["grad/" + k] = g.numpy()
            else:
                arrs["gsample/" + k] = sample(g)
                arrs["gsum/" + k] = np.array([g.double().sum().item(), g.double().abs().sum().item()])
    # routing artefacts: re-execute core.py:480-492,529 on the same inputs (ops of the reference's own module)
    with torch.no_grad():
        ffn = layer.feed_forward.ffn
        h = layer.attention(x.detach())[0]
        n2 = layer.feed_forward.pre_norm(h).reshape(-1, spec["Dm"])
        logits = ffn.router(ffn.router_norm(n2)).float()
        if training:
            logits = logits + noise * (torch.nn.functional.softplus(ffn.w_noise) * ffn.noisy_routing_alpha)
        gates = torch.softmax(logits, dim=-1)
        probs, idx = torch.topk(gates, spec["K"], dim=-1)
        w = probs / (probs.sum(-1, keepdim=True) + 1e-6)
        arrs.update(moe_in=n2.numpy(), logits=logits.numpy(), idx=idx.numpy(), w=w.numpy(), ssm_out=h.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), spec=json.dumps(spec), **arrs)
    print(name, "out", arrs["out"].shape, "lb", float(arrs["lb"]), "rz", float(arrs["rz"]))


def case_ssm_cache(name, spec):
    """Prefill with use_cache, then single-token decode steps through past_key_value (core.py:363-372,391-400)."""
    layer, sd = ref_layer(spec)
    layer.eval()
    ssm = layer.attention.attention_mechanism_impl
    x, _ = O.make_inputs(spec["B"], spec["L"] + spec["steps"], spec["Dm"], spec["E"], seed=spec["seed"])
    arrs = {}
    with torch.no_grad():
        full, yfull, _ = ssm(x, output_attentions=True, use_cache=False)
        arrs["full_out"], arrs["full_y"] = full.numpy(), yfull.numpy()
        out, y, cache = ssm(x[:, :spec["L"]], output_attentions=True, use_cache=True)
        arrs["prefill_out"], arrs["prefill_y"] = out.numpy(), y.numpy()
        arrs["prefill_conv"], arrs["prefill_h"] = cache[0].numpy(), cache[1].numpy()
        for s in range(spec["steps"]):
            out, y, cache = ssm(x[:, spec["L"] + s: spec["L"] + s + 1], past_key_value=cache,
                                output_attentions=True, use_cache=True)
            arrs[f"step{s}_out"], arrs[f"step{s}_conv"], arrs[f"step{s}_h"] = out.numpy(), cache[0].numpy(), cache[1].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), spec=json.dumps(spec), **arrs)
    print(name, "ok")


def case_ssm_scans(name, spec):
    """Training-mode SSM alone: log-cumsum scan (train) vs recurrent scan (eval) + grads of the train path."""
    layer, sd = ref_layer(spec)
    ssm = layer.attention.attention_mechanism_impl
    x, _ = O.make_inputs(spec["B"], spec["L"], spec["Dm"], spec["E"], seed=spec["seed"])
    x = x.requires_grad_(True)
    ssm.train()
    out, y, _ = ssm(x, output_attentions=True)
    (out.pow(2).mean() + y.pow(2).mean()).backward()
    arrs = {"train_out": out.detach().numpy(), "train_y": y.detach().numpy(), "dx": x.grad.numpy()}
    for k, p in ssm.named_parameters():
        arrs["grad/" + k] = p.grad.numpy()
    ssm.eval()
    with torch.no_grad():
        out2, y2, _ = ssm(x.detach(), output_attentions=True)
    arrs["eval_out"], arrs["eval_y"] = out2.numpy(), y2.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), spec=json.dumps(spec), **arrs)
    print(name, "train/eval max diff", float((out.detach() - out2).abs().max()))


if __name__ == "__main__":
    base = dict(Dm=64, H=2, I=128, E=8, K=2, B=2, L=48, seed=1)
    case_block("block_small_train", base)
    case_block("block_small_eval", dict(base, training=False, seed=2))
    case_block("block_relu_e4", dict(base, E=4, act="relu", L=37, seed=3))
    case_block("block_drop_expert", dict(base, E=10, seed=4, perm=[3, 7, 0, 1, 2, 4, 5, 6, 8, 9]))
    case_block("block_h3_ragged", dict(Dm=96, H=3, I=160, E=8, K=2, B=3, L=29, seed=5))
    case_block("block_c1dims", dict(Dm=256, H=4, I=1024, E=8, K=2, B=1, L=192, seed=6), full_grads=False)
    case_ssm_cache("ssm_cache_decode", dict(base, L=12, steps=3, seed=7))
    case_ssm_scans("ssm_scans_l512", dict(base, B=1, L=512, seed=8))
