"""TEST INFRASTRUCTURE ONLY (see oracle/shim/timm/__init__.py): the one symbol
/root/reference/models/moe/origin/vision_transformer_moe.py:13 imports."""
from timm.layers import lecun_normal_  # noqa: F401
