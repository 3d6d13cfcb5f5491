"""Multi-GPU plumbing: one process per GPU (torchrun), groupings sharded across ranks.

The groupings handed to SetCoverFilter are independent set-cover instances
(filter/set_cover_filter.py:817-846); the reference solves them in a process pool, biggest first
(:874-895).  Here every rank keeps the full (host) input, runs the device path only for the
groupings it owns, and the per-grouping selections (a few KB of indices) are exchanged at the end
with one all_gather over torch.distributed -- no collective on the data path.  With a single
process (no RANK in the environment) everything below is a no-op.
"""
import os
import sys


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def assign_groups(sizes, world_size):
    """Owner rank of every grouping: longest-processing-time first (the reference starts the
    largest groupings first, set_cover_filter.py:876-890): groupings in descending size, each to
    the currently least-loaded rank; ties go to the lower rank / lower index, so every rank
    computes the same assignment."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world_size
    owner = [0] * len(sizes)
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += max(int(sizes[i]), 1)
    return owner


def _dist():
    # A process group can only be initialised if the launcher already imported torch.distributed;
    # never import torch from here (seconds of start-up a single-GPU run does not need).
    dist = sys.modules.get('torch.distributed')
    if dist is None:
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def exchange_group_results(local, owner, rank):
    """local: {group index: result} for the groupings this rank owns.  Returns the list of all
    results in grouping order on every rank."""
    n = len(owner)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return [local[i] for i in range(n)]
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, local)
    merged = {}
    for part in gathered:
        merged.update(part)
    missing = [i for i in range(n) if i not in merged]
    if missing:
        raise RuntimeError("groupings %s were not computed by any rank" % missing)
    return [merged[i] for i in range(n)]


def active():
    """True when running under an initialised multi-rank process group."""
    dist = _dist()
    return dist is not None and dist.get_world_size() > 1


def shard_bounds(n, world_size, rank):
    """Contiguous block [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def ensure_comm(ctx):
    """Join the NCCL communicator of libcatchb200 (once per context): rank 0 creates the unique
    id, torch.distributed carries the 128 bytes to the other ranks."""
    if getattr(ctx, 'comm_ready', False):
        return
    dist = _dist()
    if dist is None:
        raise RuntimeError("probe sharding needs an initialised torch.distributed process group")
    rank, world_size = dist.get_rank(), dist.get_world_size()
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world_size)
