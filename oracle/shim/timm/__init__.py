"""TEST INFRASTRUCTURE ONLY.  Stub for the one timm symbol the reference's
vision_transformer_moe.py imports at module scope (`timm.layers.lecun_normal_`)."""
