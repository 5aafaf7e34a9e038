"""TEST INFRASTRUCTURE ONLY (see oracle/shim/timm/__init__.py)."""
import math
import torch


def lecun_normal_(tensor):
    fan_in = tensor.shape[1] if tensor.dim() > 1 else tensor.shape[0]
    with torch.no_grad():
        return tensor.normal_(0, math.sqrt(1.0 / fan_in))
