"""TEST INFRASTRUCTURE ONLY.  Restatement of fmoe/functions.py (FastMoE @4edeccd)
for world_size == 1, in plain PyTorch (see oracle/shim/fmoe/__init__.py)."""
import torch
from torch.autograd import Function


def ensure_comm(t, comm):  # upstream: lazily builds fmoe's ncclComm_t
    return None


def count_by_gate(gate, num_expert, world_size, require_pos=True):
    """upstream fmoe/functions.py::count_by_gate -- expert_count + assign_pos."""
    with torch.no_grad():
        flat = gate.reshape(-1)
        local_expert_count = torch.bincount(
            flat[flat >= 0], minlength=num_expert * world_size
        ).to(torch.int32)[: num_expert * world_size]
        global_expert_count = local_expert_count.clone()  # world_size == 1
        if not require_pos:
            pos = None
        else:
            # stable: rows of one expert keep flat-slot order
            pos = torch.sort(flat, stable=True).indices
    return pos, local_expert_count, global_expert_count


def prepare_forward(gate, num_expert, world_size):
    pos, local_expert_count, global_expert_count = count_by_gate(gate, num_expert, world_size)
    with torch.no_grad():
        fwd_expert_count = global_expert_count.view(world_size, num_expert).sum(dim=0)
        fwd_batch_size = int(fwd_expert_count.sum().item())
    return (
        pos,
        local_expert_count.cpu(),
        global_expert_count.cpu(),
        fwd_expert_count.cpu(),
        fwd_batch_size,
    )


def _local_scatter(inp, pos):
    return torch.index_select(inp, 0, pos)


def _local_gather(inp, pos, out_batch_size, maybe_overlap=True):
    inp_buf = torch.zeros(out_batch_size, inp.shape[-1], dtype=inp.dtype, device=inp.device)
    if maybe_overlap:
        inp_buf.index_add_(0, pos, inp)
    else:
        inp_buf.index_copy_(0, pos, inp)
    return inp_buf


class MOEScatter(Function):
    @staticmethod
    def forward(ctx, inp, pos, local_expert_count, global_expert_count, fwd_batch_size, world_size):
        local_input_buf = _local_scatter(inp, pos)
        ctx.moe_args = inp.shape[0], pos.shape[0], world_size
        ctx.save_for_backward(pos)
        return local_input_buf

    @staticmethod
    def backward(ctx, global_grad_in):
        (pos,) = ctx.saved_tensors
        inp_batch_size, _, _ = ctx.moe_args
        grad_in = _local_gather(global_grad_in, pos, inp_batch_size)
        return grad_in, None, None, None, None, None


class MOEGather(Function):
    @staticmethod
    def forward(ctx, global_output_buf, pos, local_expert_count, global_expert_count,
                local_batch_size, world_size):
        output = _local_gather(global_output_buf, pos, local_batch_size, maybe_overlap=False)
        ctx.save_for_backward(pos)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        (pos,) = ctx.saved_tensors
        return _local_scatter(grad_out.contiguous(), pos), None, None, None, None, None


class MOELinear(Function):
    """Per-expert `x_e @ W[e].T + b[e]` over contiguous queue segments
    (upstream: one cuBLAS GEMM per expert, cuda/parallel_linear.cuh)."""

    @staticmethod
    def forward(ctx, global_input_buf, fwd_expert_count, weight, bias=None):
        out = global_input_buf.new_empty(global_input_buf.shape[0], weight.shape[1])
        base = 0
        for e, n in enumerate(fwd_expert_count.tolist()):
            if n:
                seg = global_input_buf[base:base + n]
                y = seg @ weight[e].t()
                if bias is not None:
                    y = y + bias[e]
                out[base:base + n] = y
            base += n
        ctx.save_for_backward(global_input_buf, fwd_expert_count, weight, bias)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        inp, cnt, weight, bias = ctx.saved_tensors
        grad_inp = torch.zeros_like(inp)
        grad_w = torch.zeros_like(weight)
        grad_b = torch.zeros_like(bias) if bias is not None else None
        base = 0
        for e, n in enumerate(cnt.tolist()):
            if n:
                g = grad_out[base:base + n]
                grad_inp[base:base + n] = g @ weight[e]
                grad_w[e] = g.t() @ inp[base:base + n]
                if grad_b is not None:
                    grad_b[e] = g.sum(0)
            base += n
        return grad_inp, None, grad_w, grad_b


class Slice(Function):  # never reached: the reference passes no slice_group
    @staticmethod
    def forward(ctx, inp, rank, world_size, group):
        raise NotImplementedError("fmoe slice parallelism is not exercised by the reference")


class AllGather(Function):
    @staticmethod
    def forward(ctx, inp, rank, world_size, group):
        raise NotImplementedError("fmoe slice parallelism is not exercised by the reference")
