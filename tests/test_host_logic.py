"""CPU-only tests of the host-side mirror of the reference interface (no kernels are launched)."""
import pytest
import torch

from apertis_llm_b200 import AdaptiveExpertSystem, ApertisLayerB200, BlockConfig, SelectiveLinearAttention, ops
from oracle import apertis_oracle as O


def test_state_dict_keys_match_reference_layout():
    cfg = BlockConfig(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_experts=4)
    layer = ApertisLayerB200(cfg)
    sd = O.make_layer_params(64, 2, 128, 4, seed=0)
    layer.load_state_dict(sd, strict=True)
    out = layer.state_dict()
    assert set(out) == set(sd)
    assert all(torch.equal(out[k], sd[k]) for k in sd)
    ffn = layer.feed_forward.ffn
    assert ffn.expert_w1.shape == (4, 128, 64) and ffn.expert_w2.shape == (4, 64, 128)
    assert torch.equal(ffn.expert_w1[2], sd["feed_forward.ffn.experts.2.1.weight"])


def test_modules_refuse_cpu_tensors():
    cfg = BlockConfig(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_experts=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SelectiveLinearAttention(cfg)(torch.randn(1, 8, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AdaptiveExpertSystem(cfg).eval()(torch.randn(1, 8, 64))


def test_passthrough_when_no_experts():
    cfg = BlockConfig(hidden_size=64, num_attention_heads=2, num_experts=0, experts_per_token=0)
    m = AdaptiveExpertSystem(cfg)
    x = torch.randn(2, 3, 64)
    out, lb, rz = m(x)                      # core.py:474-475
    assert out is x and float(lb) == 0.0 and float(rz) == 0.0


def test_capacity_formula():
    assert ops.moe_capacity(4096, 8, 1.25, True, True) == 640      # core.py:510
    assert ops.moe_capacity(4096, 8, 1.25, False, True) == 4096
    assert ops.moe_capacity(3, 8, 1.25, True, True) == 1
    assert ops.moe_capacity(96, 8, 1.25, True, True) == O.moe_capacity(96, 8)


def test_config_derivations():
    cfg = BlockConfig(hidden_size=704, num_attention_heads=11)
    assert cfg.ssm_d_inner == 176 and cfg.ssm_dt_rank == 44
    m = SelectiveLinearAttention(cfg)
    assert m.x_param_proj.weight.shape == (44 + 2 * 176, 176) and m.A_log.shape == (11, 16)
