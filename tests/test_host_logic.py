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


def test_state_dict_rejects_malformed_expert_keys():
    """Reference checkpoints load with strict=True; a key under 'experts.' that is not a well-formed key of an existing
    expert is reported as unexpected instead of being dropped silently."""
    cfg = BlockConfig(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_experts=4)
    m = AdaptiveExpertSystem(cfg)
    sd = m.state_dict()
    assert "experts.3.4.bias" in sd and not any(k.startswith("expert_") for k in sd)      # reference key names only
    AdaptiveExpertSystem(cfg).load_state_dict(sd, strict=True)
    for bad in ("experts.7.1.weight", "experts.1.9.weight", "experts.x.1.weight"):
        sd2 = dict(sd)
        sd2[bad] = torch.zeros(1)
        with pytest.raises(RuntimeError, match="Unexpected key"):
            AdaptiveExpertSystem(cfg).load_state_dict(sd2, strict=True)


def test_kernel_limits_are_checked_at_construction():
    with pytest.raises(ValueError, match="num_experts <= 32"):
        AdaptiveExpertSystem(BlockConfig(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_experts=40))
    with pytest.raises(ValueError, match="multiples of 8"):
        AdaptiveExpertSystem(BlockConfig(hidden_size=60, num_attention_heads=2, intermediate_size=128, num_experts=4))
    from apertis_llm_b200 import ApertisLayerB200
    with pytest.raises(NotImplementedError, match="use_rmsnorm"):
        ApertisLayerB200(BlockConfig(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_experts=4, use_rmsnorm=True))
