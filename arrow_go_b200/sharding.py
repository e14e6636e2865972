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


def global_min_max(local_min, local_max, dist, device="cpu"):
    """min/max of the per-rank (min, max) pairs (ag_min_max per GPU + two 8-byte all-reduces).  Values
    travel as int64; uint64 callers pass their values through int(np.int64(np.uint64(v))) ^ (1 << 63) style
    order-preserving maps if they exceed 2^63 (not needed for the Parquet physical types, which are signed)."""
    import torch
    lo = torch.tensor([int(local_min)], dtype=torch.int64, device=device)
    hi = torch.tensor([int(local_max)], dtype=torch.int64, device=device)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return int(lo.item()), int(hi.item())


def cumulative_sum_carry(local_total, local_has_null, dist, skip_nulls=False, device="cpu"):
    """What a row-range-sharded cumulative_sum needs from the other ranks: the wrapping int64 sum of every
    LOWER rank's shard (its start value: pass it to ag_cumulative_sum_state_init_dev) and whether a lower
    rank met a null (without skip_nulls this rank's whole output is then null).  Each rank first runs the
    scan once with start 0 (or just Sum over its valid rows up to its first null) to get `local_total`;
    G x 9 bytes cross the wire."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([np.int64(np.uint64(local_total % (1 << 64)).astype(np.int64)), int(bool(local_has_null))], dtype=torch.int64, device=device)
    parts = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(parts, mine)
    start, dead = 0, False
    for r in range(rank):
        if not skip_nulls and dead:
            break
        start = (start + int(parts[r][0].item())) % (1 << 64)
        dead = dead or bool(parts[r][1].item())
    if start >= 1 << 63:
        start -= 1 << 64
    return start, (dead and not skip_nulls)


# ---------------------------------------------------------------------------------------------------------
# The product's own communicator (include/arrowgpu.h "Multi-GPU"): HBM mailboxes written over NVLink by the Sum kernel
# itself.  torch.distributed is only the out-of-band channel that all-gathers the 64-byte IPC handles at start-up.
def create_comm(dist=None, device="cpu"):
    """One rank's ag_comm_t.  world 1 (dist None): a self-contained communicator.  Otherwise every rank exports its
    mailbox handle, the handles are all-gathered through `dist` (any backend), and the peers' mailboxes are mapped."""
    import ctypes as C

    from . import _native as N
    comm = C.c_void_p()
    if dist is None or dist.get_world_size() == 1:
        N.call("ag_comm_create", C.byref(comm), 1, 0, None)
        return comm
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    h = (C.c_ubyte * 64)()
    N.call("ag_comm_local_handle", world, h)
    mine = torch.tensor(list(bytes(h)), dtype=torch.uint8, device=device)
    parts = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(parts, mine)
    table = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
    buf = (C.c_ubyte * (64 * world)).from_buffer_copy(table)
    N.call("ag_comm_create", C.byref(comm), world, rank, buf)
    return comm


def attach_nccl(comm, dist, device="cpu"):
    """Optional: give the communicator an NCCL handle (rank 0 draws the id, `dist` broadcasts its 128 bytes)."""
    import ctypes as C

    import torch

    from . import _native as N
    ident = (C.c_ubyte * 128)()
    if dist.get_rank() == 0:
        N.call("ag_comm_unique_id", ident)
    t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    ident = (C.c_ubyte * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
    N.call("ag_comm_attach_nccl", comm, ident)
