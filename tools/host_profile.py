"""cProfile of the host side of one layer call at a tiny batch (launch-bound): where do the microseconds go?"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import m3vit_b200 as M
from m3vit_b200.synthetic import device_tokens

dev = torch.device("cuda:0")
D = H = 384
layer = M.FMoETransformerMLP(num_expert=16, d_model=D, d_gate=D + 2, d_hidden=H,
                             activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE, top_k=4,
                             vmoe_noisy_std=0, multi_gate=True, compute_dtype=torch.bfloat16).to(dev).train()
T = 2402
x = device_tokens(T, D, 0, dev).requires_grad_(True)
g = torch.randn(T, D, device=dev) * 0.01
w = torch.tensor(0.01, device=dev)


def fwd():
    return layer(x, task_id=0)


def step():
    x.grad = None
    out = fwd()
    torch.autograd.backward([out, layer.gate[0].get_loss()], [g, w])


for _ in range(20):
    step()
torch.cuda.synchronize()
for name, fn in (("forward only", fwd), ("fwd+bwd", step)):
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name}: host {1e6 * (t1 - t0) / 200:.1f} us per call")
prof = cProfile.Profile()
prof.enable()
for _ in range(200):
    fwd()
prof.disable()
torch.cuda.synchronize()
pstats.Stats(prof).sort_stats("tottime").print_stats(22)
