"""Multi-GPU sharding of independent planning problems.

The reference's eval loop (`for env_idx, env_config in enumerate(...)`, eval_planning_3d.py:101-126)
never shares state between problems, so the path shards with NO data-path collective: every rank
(one process per GPU) plans a contiguous slice of the problem list, and the per-problem result rows
(`path_len_list`, padded) are gathered ONCE at the end -- `all_gather` over NCCL on the GPUs, over
gloo in the CPU tests.
"""
import numpy as np


def shard_bounds(n_items, world, rank):
    """Contiguous slice [begin, end) of rank `rank`; the first n_items % world ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pack_rows(lists, width=None):
    """Variable-length per-problem lists -> (rows [n][width] f64 padded with NaN, lengths [n] int64)."""
    n = len(lists)
    width = max([len(l) for l in lists] + [1]) if width is None else width
    rows = np.full((n, width), np.nan)
    lens = np.zeros(n, dtype=np.int64)
    for i, l in enumerate(lists):
        if len(l) > width:
            raise ValueError("row longer than the padded width")
        rows[i, :len(l)] = l
        lens[i] = len(l)
    return rows, lens


def unpack_rows(rows, lens):
    return [list(rows[i, :int(lens[i])]) for i in range(len(lens))]


def gather_lists(local_lists, n_total, device=None):
    """Every rank passes the result lists of ITS slice (shard_bounds order); every rank receives the
    n_total lists in global problem order.  One all_gather of a fixed-shape padded tensor (+ one of
    the lengths); without an initialised process group it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(l) for l in local_lists]
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-n_total // world)                                   # rows per rank after padding
    width = torch.tensor([max([len(l) for l in local_lists] + [1])], dtype=torch.int64, device=device)
    dist.all_reduce(width, op=dist.ReduceOp.MAX)
    rows, lens = pack_rows(local_lists, int(width.item()))
    pad = per - len(local_lists)
    rows = np.concatenate([rows, np.full((pad, rows.shape[1]), np.nan)]) if pad else rows
    lens = np.concatenate([lens, np.full(pad, -1, dtype=np.int64)]) if pad else lens
    t_rows = torch.from_numpy(rows).to(device) if device is not None else torch.from_numpy(rows)
    t_lens = torch.from_numpy(lens).to(device) if device is not None else torch.from_numpy(lens)
    g_rows = [torch.empty_like(t_rows) for _ in range(world)]
    g_lens = [torch.empty_like(t_lens) for _ in range(world)]
    dist.all_gather(g_rows, t_rows)
    dist.all_gather(g_lens, t_lens)
    out = []
    for r in range(world):
        b, e = shard_bounds(n_total, world, r)
        out += unpack_rows(g_rows[r].cpu().numpy()[:e - b], g_lens[r].cpu().numpy()[:e - b])
    return out
