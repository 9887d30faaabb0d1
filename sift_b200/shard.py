"""Data-parallel plumbing: images are independent units, so a batch is split into
contiguous per-rank ranges and nothing is exchanged on the data path.  torch.distributed is used only
for the barrier and for reducing the timing (max over ranks) and the counters (sum)."""
import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items, rank, world_size):
    """Contiguous range [lo, hi) of the items owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_process_group(backend):
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return dist


def reduce_max_sum(dist, device, max_vals, sum_vals):
    """All-reduce: element-wise max of max_vals and sum of sum_vals (lists of floats) over all ranks."""
    import torch

    m = torch.tensor(list(max_vals), dtype=torch.float64, device=device)
    s = torch.tensor(list(sum_vals), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return m.tolist(), s.tolist()
