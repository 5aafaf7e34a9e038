"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/moe_oracle.py header).

Expert-parallel exchange plan, restating what FastMoE's global_scatter produces on the
receiving rank (fmoe_cuda global_exchange: receive buffer laid out expert-major then
source-rank) for the sharding rule of the reference (utils/common_config.py:179-185:
expert e lives on rank e // E_loc), with the B200 padded layout (each local expert's queue
rounded up to `pad` rows)."""
import torch

from .moe_oracle import route_plan


def ep_plan(idx, cnt_all, rank, W, E_loc, pad):
    """idx [T,K] int64 global expert ids of THIS rank's slots; cnt_all [W, W*E_loc] int.
    returns dst_rank[T*K], dst_row[T*K], recv_counts[E_loc], recv_offsets[E_loc+1] (for `rank`)"""
    E_tot = W * E_loc
    cnt_all = cnt_all.long()
    flat = idx.reshape(-1).long()
    _, off_local, pos_local, _ = route_plan(idx, E_tot, 1)       # stable local order
    tot = cnt_all.sum(0)                                         # rows per global expert
    before = cnt_all[:rank].sum(0)                               # rows from lower ranks
    base = torch.zeros(E_tot, dtype=torch.int64)
    for o in range(W):
        roff = 0
        for le in range(E_loc):
            ge = o * E_loc + le
            base[ge] = roff + before[ge]
            roff += (int(tot[ge]) + pad - 1) // pad * pad
    dst_rank = flat // E_loc
    dst_row = base[flat] + (pos_local.long() - off_local.long()[flat])
    mine = tot[rank * E_loc:(rank + 1) * E_loc]
    padded = (mine + pad - 1) // pad * pad
    recv_offsets = torch.zeros(E_loc + 1, dtype=torch.int64)
    recv_offsets[1:] = torch.cumsum(padded, 0)
    return dst_rank.int(), dst_row.int(), mine.int(), recv_offsets.int()
