"""Small-batch (launch-bound) layer call under CUDA graphs: torch.cuda.make_graphed_callables captures the forward and
the backward of the drop-in layer (no host sync anywhere in the path), replay costs two graph launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import m3vit_b200 as M
from m3vit_b200.synthetic import device_tokens


class TaskCall(nn.Module):
    """layer(x, task_id=t) -> (out, cv_loss) with tensor-only positional arguments (what graph capture needs)."""

    def __init__(self, layer, task):
        super().__init__()
        self.layer, self.task = layer, task

    def forward(self, x):
        out = self.layer(x, task_id=self.task)
        return out, self.layer.gate[self.task].get_loss()


def main():
    dev = torch.device("cuda:0")
    print("M3_KNOBS =", os.environ.get("M3_KNOBS", "(defaults)"))
    D = H = 384
    for B in ([int(a) for a in sys.argv[1:]] or [2, 8]):
        T = B * 1201
        torch.manual_seed(0)
        layer = M.FMoETransformerMLP(num_expert=16, d_model=D, d_gate=D + 2, d_hidden=H,
                                     activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE,
                                     top_k=4, vmoe_noisy_std=0, multi_gate=True, compute_dtype=torch.bfloat16).to(dev).train()
        eager = TaskCall(layer, 0)
        x = device_tokens(T, D, 0, dev).requires_grad_(True)
        g = torch.randn(T, D, device=dev) * 0.01
        w = torch.tensor(0.01, device=dev)

        def step(mod):
            x.grad = None
            out, loss = mod(x)
            torch.autograd.backward([out, loss], [g, w])
            return out

        def timeit(mod, n=50):
            for _ in range(5):
                step(mod)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                step(mod)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1000
        t_eager = timeit(eager)
        layer.zero_grad(set_to_none=True)
        out_e = step(eager).detach().clone()
        dx_e = x.grad.clone()
        dw_e = layer.experts.htoh4.weight.grad.clone()
        layer.zero_grad(set_to_none=True)
        graphed = torch.cuda.make_graphed_callables(TaskCall(layer, 0), (x.detach().clone().requires_grad_(True),),
                                                     allow_unused_input=True)   # the other task gates get no grad
        layer.zero_grad(set_to_none=True)
        out_g = step(graphed).detach().clone()
        ok = torch.equal(out_g, out_e) and torch.equal(x.grad, dx_e) and torch.equal(layer.experts.htoh4.weight.grad, dw_e)
        t_graph = timeit(graphed)
        print(f"B={B} T={T}: eager {t_eager:7.1f} us/call  graphed {t_graph:7.1f} us/call  "
              f"({T / t_graph:.1f} Mtok/s)  bit-identical to eager: {ok}", flush=True)


if __name__ == "__main__":
    main()
