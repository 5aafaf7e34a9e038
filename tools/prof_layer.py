"""One MoE layer call fwd+bwd at bench size (for ncu / launch lists).  python tools/prof_layer.py [B] [dtype]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cdt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "fp32") else torch.bfloat16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
layers = bench.build_layers(dev, cdt)[:1]
from m3vit_b200.synthetic import device_tokens

T = B * bench.N_TOK
x = device_tokens(T, bench.D_MODEL, 0, dev).requires_grad_(True)
g = torch.randn(T, bench.D_MODEL, device=dev) * 0.01
for i in range(iters):
    x.grad = None
    bench.one_call(layers[0], x, g, i % 2)
torch.cuda.synchronize()
print("ok", T)
