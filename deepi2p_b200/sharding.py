"""Multi-GPU sharding of an evaluation batch.

Every (sample, init) solve is independent, so the batch of samples is block-partitioned across
ranks (one process per GPU), all inits of a sample stay on one GPU, and the only communication is
one all-gather of the final [S_local, 17] f64 records (4x4 pose row-major + cost) -- NCCL over
NVLink on GPUs, gloo in the CPU tests.  The reference has no distributed code at all
(registration_lsq.py:142-186 forks 8 OS processes on one host).
"""
import torch
import torch.distributed as dist


def partition(n_items, world_size, rank):
    """Contiguous block partition; the first n_items % world_size ranks get one extra item.
    Returns (start, stop)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_records(P, cost):
    """[S,4,4] poses + [S] costs -> [S,17] f64 records."""
    S = P.shape[0]
    return torch.cat([P.reshape(S, 16).to(torch.float64), cost.reshape(S, 1).to(torch.float64)], dim=1).contiguous()


def unpack_records(rec):
    return rec[:, :16].reshape(-1, 4, 4), rec[:, 16]


def gather_poses(P, cost, n_total=None, group=None, out=None):
    """All-gather the per-rank results into global sample order.  Returns (P [S_total,4,4], cost [S_total]) on every
    rank.

    Fast path (the benchmark's and any block partition's case): when `n_total` is given, every rank can work out all
    shard sizes from partition() alone, so there is NO size exchange and no host synchronisation -- one
    all_gather_into_tensor of the [S_local,17] f64 records (equal shards) or of records padded to the largest shard
    (uneven ones).  Without `n_total` the shard sizes are exchanged first (one extra small collective + a host read).
    `out`: optional preallocated [world * max_shard, 17] f64 buffer for the gathered records."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return P, cost
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rec = pack_records(P, cost)
    if n_total is not None:
        counts = [partition(n_total, world, r) for r in range(world)]
        counts = [b - a for a, b in counts]
        if counts[rank] != rec.shape[0]:
            raise RuntimeError("rank %d holds %d records, the block partition of %d gives it %d" % (
                rank, rec.shape[0], n_total, counts[rank]))
    else:
        n_local = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
        sizes = torch.empty((world,), dtype=torch.int64, device=rec.device)
        dist.all_gather_into_tensor(sizes, n_local, group=group)
        counts = [int(c) for c in sizes.tolist()]
    m = max(counts)
    if rec.shape[0] != m:
        padded = torch.zeros((m, 17), dtype=torch.float64, device=rec.device)
        padded[:rec.shape[0]] = rec
        rec = padded
    if out is None or tuple(out.shape) != (world * m, 17) or out.device != rec.device:
        out = torch.empty((world * m, 17), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    if min(counts) == m:
        full = out
    else:
        full = torch.cat([out[r * m:r * m + counts[r]] for r in range(world)], dim=0)
    if n_total is not None and full.shape[0] != n_total:
        raise RuntimeError("gathered %d records, expected %d" % (full.shape[0], n_total))
    return unpack_records(full)
