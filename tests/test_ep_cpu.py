"""CPU: expert-parallel host logic.  (1) the EP plan oracle is a bijection onto the owners'
queues with the expert-major / source-rank-major order; (2) world_size-2 gloo run of the
count exchange + barrier plumbing used by m3vit_b200.ep (TorchDistGroup); (3) arena allocator
determinism."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import ep_oracle


def make_rank_idx(rank, T, K, E_tot, seed=0):
    gen = torch.Generator().manual_seed(seed * 100 + rank)
    return torch.stack([torch.randperm(E_tot, generator=gen)[:K] for _ in range(T)])


def check_global_plan(W, E_loc, K, Ts, pad):
    E_tot = W * E_loc
    idxs = [make_rank_idx(r, Ts[r], K, E_tot) for r in range(W)]
    cnt_all = torch.stack([torch.bincount(i.reshape(-1), minlength=E_tot) for i in idxs])
    plans = [ep_oracle.ep_plan(idxs[r], cnt_all, r, W, E_loc, pad) for r in range(W)]
    for o in range(W):                                   # owner o: who lands where
        rc, ro = plans[o][2], plans[o][3]
        rows, src, exp, slot = [], [], [], []
        for r in range(W):
            dr, drow = plans[r][0], plans[r][1]
            m = dr == o
            rows.append(drow[m]); src.append(torch.full((int(m.sum()),), r))
            exp.append(idxs[r].reshape(-1)[m] - o * E_loc); slot.append(m.nonzero().flatten())
        rows, src, exp, slot = map(torch.cat, (rows, src, exp, slot))
        assert rows.unique().numel() == rows.numel(), "two slots collide on one queue row"
        assert int(rc.sum()) == rows.numel()
        for le in range(E_loc):
            m = exp == le
            r_le = rows[m].long()
            assert torch.equal(r_le.sort().values, torch.arange(int(rc[le])) + int(ro[le]))      # contiguous
            order = r_le.argsort()
            s_sorted, slot_sorted = src[m][order], slot[m][order]
            assert torch.equal(s_sorted, s_sorted.sort().values), "rows of an expert must be source-rank-major"
            for r in range(W):
                ss = slot_sorted[s_sorted == r]
                assert torch.equal(ss, ss.sort().values), "rows of one source keep slot order (stable)"
        assert all(int(v) % pad == 0 for v in ro)


@pytest.mark.parametrize("W,E_loc,K,pad", [(2, 8, 4, 128), (4, 4, 2, 128), (8, 2, 4, 128), (2, 2, 1, 1)])
def test_ep_plan_oracle_is_consistent(W, E_loc, K, pad):
    check_global_plan(W, E_loc, K, [50 + 7 * r for r in range(W)], pad)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from m3vit_b200.ep import TorchDistGroup
    grp = TorchDistGroup()
    E_loc, K, T = 4, 2, 40 + rank
    E_tot = world * E_loc
    idx = make_rank_idx(rank, T, K, E_tot)
    counts = torch.bincount(idx.reshape(-1), minlength=E_tot).int()
    cnt_all = grp.all_gather_counts(counts)
    grp.barrier(torch.device("cpu"))
    handles = grp.exchange_bytes(bytes([rank]) * 64)
    dr, drow, rc, ro = ep_oracle.ep_plan(idx, cnt_all, rank, world, E_loc, 128)
    gathered = [None] * world
    dist.all_gather_object(gathered, (dr, drow, rc, ro, counts))
    ok = True
    ok &= all(torch.equal(cnt_all[r], gathered[r][4]) for r in range(world))
    ok &= [h[0] for h in handles] == list(range(world))
    # rows landing on me are exactly my recv_counts, without collisions
    mine = torch.cat([g[1][g[0] == rank] for g in gathered])
    ok &= mine.unique().numel() == mine.numel() == int(rc.sum())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_count_exchange_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_arena_allocator_is_deterministic():
    """Two ranks replaying the same alloc/free sequence must get identical offsets (no GPU needed:
    exercise the allocator logic on an Arena shell)."""
    from m3vit_b200.ep import Arena

    def replay():
        a = Arena.__new__(Arena)
        a.nbytes, a._top, a._free = 1 << 30, 0, {}
        offs = []
        x = a.alloc(1000); y = a.alloc(5000); offs += [x, y]
        a.free(x, 1000)
        z = a.alloc(900); offs.append(z)          # same size class (1024) -> reuses x
        w = a.alloc(5000); offs.append(w)
        a.free(y, 5000); a.free(w, 5000)
        offs.append(a.alloc(4097))
        return offs
    assert replay() == replay()
    o = replay()
    assert o[2] == o[0] and all(v % 1024 == 0 for v in o)


def _sync_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn as nn
    import m3vit_b200 as M
    from m3vit_b200 import dist_utils as U
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(10 + rank)                       # every rank starts with DIFFERENT weights
    blk = nn.Module()
    blk.mlp = M.FMoETransformerMLP(num_expert=2, d_model=16, d_gate=16 + 2, d_hidden=16,
                                   activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE, world_size=world,
                                   top_k=2, vmoe_noisy_std=0, multi_gate=True)
    blk.norm = nn.LayerNorm(16)
    with torch.no_grad():
        blk.norm.weight.add_(rank + 1.0)
    before = {k: v.clone() for k, v in blk.state_dict().items()}
    U.sync_weights(blk)
    after = blk.state_dict()
    gathered = {}
    ok = True
    for k, v in after.items():
        both = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(both, v.contiguous())
        expert = any(w in k for w in U.EXPERT_KEY_WORDS)
        if expert:
            ok &= torch.equal(v, before[k])                                           # rank-local, untouched
            if k.endswith("weight"):                                                  # (biases start at zero everywhere)
                ok &= not torch.equal(both[0], both[1])
        else:
            ok &= torch.equal(both[0], both[1])                                       # rank 0's copy everywhere
            if rank == 0:
                ok &= torch.equal(v, before[k])
    # gradients: the router and the LayerNorm average over the world, experts keep theirs, an unused task gate gets zeros
    tags = {n: getattr(p, "dp_comm", "dp") for n, p in blk.named_parameters()}
    ok &= tags["mlp.experts.htoh4.weight"] == "none" and tags["mlp.gate.0.w_gate"] == "gate"
    for n, p in blk.named_parameters():
        p.grad = None if n == "mlp.gate.1.w_gate" else torch.full_like(p, float(rank + 1))
    nred = U.allreduce_replicated_grads(blk)
    ok &= nred == 4                                                                   # 2 task gates + LayerNorm weight, bias
    for n, p in blk.named_parameters():
        if tags[n] == "none":
            ok &= bool((p.grad == rank + 1).all())
        elif n == "mlp.gate.1.w_gate":
            ok &= bool((p.grad == 0).all())
        else:
            ok &= bool((p.grad == 1.5).all())                                         # mean of 1 and 2
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_replicated_parameter_sync_world2_gloo():
    """sync_weights / allreduce_replicated_grads (the reference's sync_weights + allreduce_params contract around an
    expert-parallel layer) on 2 gloo ranks: experts stay rank-local, everything else follows rank 0 / averages."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
