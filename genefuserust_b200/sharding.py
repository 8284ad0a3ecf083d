"""Multi-GPU sharding of a batch: pairs are independent units (map_read(&self), fusion_mapper.rs:93), so rank r
maps a contiguous range against its own replica of the index and the (tiny) match records are gathered on the
host.  No data-path collective; torch.distributed is used only for the control-plane gather/barrier."""
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous [lo, hi) of rank; sizes differ by at most one pair"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def map_shard(map_fn, batch, rank, world):
    """map_fn(sub_batch) -> list of record tuples whose first field is the LOCAL pair index.
    Returns records with GLOBAL pair indices for this rank's shard."""
    lo, hi = shard_range(batch.n, rank, world)
    recs = map_fn(batch.slice(lo, hi)) if hi > lo else []
    return [(r[0] + lo,) + tuple(r[1:]) for r in recs]


def gather_matches(local_records, group=None):
    """all ranks receive the concatenation of every rank's records, ordered by (pair_idx, source) — the order
    gf_map_pairs itself returns for an unsharded batch."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sorted(local_records, key=lambda r: (r[0], r[1]))
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local_records, group=group)
    out = [r for p in parts for r in p]
    out.sort(key=lambda r: (r[0], r[1]))
    return out
