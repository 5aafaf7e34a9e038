"""TEST INFRASTRUCTURE ONLY.  Restatement of fmoe/linear.py (FastMoE @4edeccd)."""
import math
import torch
import torch.nn as nn
from .functions import MOELinear


class FMoELinear(nn.Module):
    """Bank of `num_expert` independent linears.  weight [E, out, in], bias [E, out]."""

    def __init__(self, num_expert, in_feat, out_feat, bias=True, rank=0):
        super().__init__()
        self.num_expert, self.in_feat, self.out_feat, self.rank = num_expert, in_feat, out_feat, rank
        self.weight = nn.Parameter(torch.Tensor(num_expert, out_feat, in_feat))
        if bias:
            self.bias = nn.Parameter(torch.zeros(num_expert, out_feat))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def forward(self, inp, fwd_expert_count):
        return MOELinear.apply(inp, fwd_expert_count, self.weight, self.bias)

    def reset_parameters(self):
        # upstream: per-expert kaiming_uniform_(a=sqrt(5)) like nn.Linear; bias zero
        for i in range(self.num_expert):
            torch.nn.init.kaiming_uniform_(self.weight[i], a=math.sqrt(5))
