"""Row-range sharding across the GPUs of one box: one process per GPU, each rank
owns a contiguous range of rows cut at multiples of 64 rows (so no two ranks share a bitmap
word), no data-path collective for Add / compare / filter / take, and one 8-byte all-reduce for
the global Sum.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only."""
import numpy as np

ALIGN_ROWS = 64


def shard_range(n_rows, rank, world, align=ALIGN_ROWS):
    """[start, stop) of `rank`: ceil-balanced, cut points are multiples of `align` (last shard takes the tail)."""
    per = -(-n_rows // world)
    per = -(-per // align) * align
    start = min(rank * per, n_rows)
    stop = min(start + per, n_rows)
    return start, stop


def all_ranges(n_rows, world, align=ALIGN_ROWS):
    return [shard_range(n_rows, r, world, align) for r in range(world)]


def global_sum_int(local_value, dist, device="cpu"):
    """Wrapping int64 sum of the per-rank partial sums (exact in any order)."""
    import torch
    t = torch.tensor([np.int64(np.uint64(local_value % (1 << 64)).astype(np.int64))], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def global_sum_float(local_value, dist, device="cpu"):
    """float64 sum with a FIXED association order (rank 0 + rank 1 + ...), so the result does not
    depend on the collective's internal reduction tree."""
    import torch
    world = dist.get_world_size()
    parts = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(parts, torch.tensor([local_value], dtype=torch.float64, device=device))
    acc = 0.0
    for p in parts:
        acc = acc + float(p.item())
    return acc


def filter_output_offsets(local_count, dist, device="cpu"):
    """Exclusive scan of the per-rank filter output lengths: where each rank's compacted slice
    lands in the global output (only G integers cross the wire)."""
    import torch
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_count], dtype=torch.int64, device=device))
    counts = [int(c.item()) for c in counts]
    return sum(counts[: dist.get_rank()]), sum(counts)
