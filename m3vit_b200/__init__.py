"""m3vit_b200 -- B200-native implementation of the M3ViT MoE-layer hot path.

Public surface (mirrors the reference's `models/moe` names):

    FMoETransformerMLP, FMoETransformerMLPCkpt   m3vit_b200.custom_moe_layer
    NoisyGate_VMoE, TokenNoisyGate_VMoE          m3vit_b200.noisy_gate_vmoe
    TokenFMoETransformerMLP                      m3vit_b200.custom_moe_layer (token-MoE experts-only entry)
    build_moe_mlp, MoEBlockMlp                   m3vit_b200.block
    ops (one wrapper per C-ABI entry point)      m3vit_b200.ops

Everything computes in hand-written CUDA behind the C ABI of include/m3vit_moe.h
(m3vit_b200/lib/libm3vit_moe.so).  Importing this package does not need a GPU;
calling anything does, and fails loudly otherwise.
"""
from .custom_moe_layer import (FMoETransformerMLP, FMoETransformerMLPCkpt, TokenFMoETransformerMLP,  # noqa: F401
                               FMoELinear)
from .noisy_gate_vmoe import NoisyGate_VMoE, TokenNoisyGate_VMoE, cv_squared  # noqa: F401
from .block import (build_moe_mlp, MoEBlockMlp, collect_noisy_gating_loss, collect_moe_activation,  # noqa: F401
                    set_moe_layer_train_mode)

__all__ = ["FMoETransformerMLP", "FMoETransformerMLPCkpt", "TokenFMoETransformerMLP", "FMoELinear", "NoisyGate_VMoE", "TokenNoisyGate_VMoE", "cv_squared",
           "build_moe_mlp", "MoEBlockMlp", "collect_noisy_gating_loss", "collect_moe_activation",
           "set_moe_layer_train_mode"]
