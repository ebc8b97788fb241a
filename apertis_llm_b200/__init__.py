"""apertis_llm_b200 -- B200-native (sm_100a) implementation of the Apertis SSM + MoE block hot path.

Public surface (mirrors the reference's module contract, src/model/core.py):
    SelectiveLinearAttention, AdaptiveExpertSystem   drop-in nn.Modules
    ApertisLayerB200                                  the block built from them
    patch_apertis_model(model)                        swap them into a reference ApertisModel in place
    BlockConfig                                       the config fields the path reads
The arithmetic lives in libapertis_b200.so (C ABI: include/apertis_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .modules import (AdaptiveExpertSystem, ApertisLayerB200, BlockConfig, SelectiveLinearAttention,  # noqa: F401
                      patch_apertis_model)

__all__ = ["SelectiveLinearAttention", "AdaptiveExpertSystem", "ApertisLayerB200", "BlockConfig", "patch_apertis_model"]
