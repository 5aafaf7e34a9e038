"""One call of the router / mover kernels at the bench size (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from m3vit_b200 import ops
from m3vit_b200._lib import PAD_ROWS

dev = torch.device("cuda:0")
T, D, E, K = 38432, 384, 16, 4
torch.manual_seed(0)
x = torch.randn(T, D, device=dev)
wg = (torch.rand(D, E, device=dev) * 2 - 1) / 4
for _ in range(2):
    g = ops.gate_fwd(x, wg, K)
    plan = ops.route_plan(g.idx, E, PAD_ROWS, g.imp_partial, g.load_partial)
    dscore = torch.randn(T, K, device=dev)
    dz, dw, _, _ = ops.gate_bwd(x, wg, g.clean_logits, g.idx_full, K, dscore=dscore)
    dxq = torch.randn(plan.cap_rows, D, device=dev).bfloat16()
    dx = ops.dispatch_bwd(dxq, plan, T, K, dz=dz, w_gate=wg)
torch.cuda.synchronize()
print("ok")
