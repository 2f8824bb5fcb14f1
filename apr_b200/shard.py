"""Multi-GPU partitioning of the hot path: pairs are independent units (datasets/dataloader.py:76 asserts one pair per
collate; InstanceNorm statistics are per pair), so rank r of W simply takes pairs r, r+W, r+2W, ... — no data-path
collective. Only the timing is reduced (MAX over ranks) to report whole-job throughput."""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Indices of the items (pairs) owned by `rank` under round-robin sharding."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def aggregate_throughput(local_items, local_seconds, device=None):
    """Whole-job throughput = sum of items over ranks / max of seconds over ranks (works with gloo or nccl)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
        n = torch.tensor([float(local_items)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        return n.item() / t.item(), n.item(), t.item()
    return local_items / local_seconds, float(local_items), float(local_seconds)
