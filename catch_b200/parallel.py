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


def exchange_group_results(local, owner, rank, failure=None, token=0):
    """local: {group index: list of selected indices} for the groupings this rank owns.  Returns the list of all
    results in grouping order on every rank.  Two fixed-shape all-reduces over the control backend (gloo,
    CPU tensors): the lengths of every grouping's result plus one failure slot per rank, then one flat
    index tensor that each rank fills in at its groupings' offsets -- no pickling, no per-object
    collectives.  `failure`: the exception that stopped this rank, if any; its presence is exchanged
    with the lengths so that EVERY rank raises instead of some of them waiting for ever in a collective.
    `token`: a checksum of the input lists (coverage.fingerprint_lists); the results are indices into those lists, so
    every rank must hold the same lists in the same order -- compared here, in the same all-reduce."""
    n = len(owner)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        if failure is not None:
            raise failure
        return [local[i] for i in range(n)]
    torch = sys.modules['torch']
    world_size = dist.get_world_size()
    head = torch.zeros(n + 2 * world_size, dtype=torch.int64)
    for i, res in local.items():
        head[i] = len(res) + 1                       # +1: "computed", so that an empty result is told from a missing one
    if failure is not None:
        head[n + rank] = 1
    head[n + world_size + rank] = int(token) & 0x3fffffffffffffff
    dist.all_reduce(head)
    failed = [r for r in range(world_size) if int(head[n + r])]
    if failure is not None:
        raise failure
    if failed:
        raise RuntimeError("rank %d failed: see that rank's log for the exception" % failed[0])
    if len(set(head[n + world_size:].tolist())) != 1:
        raise RuntimeError("ranks hold different probe lists (or the same probes in a different order): the selected "
                           "indices cannot be exchanged.  A list ordered by a Python set differs between processes "
                           "unless PYTHONHASHSEED is pinned")
    counts = head[:n].tolist()
    missing = [i for i in range(n) if counts[i] == 0]
    if missing:
        raise RuntimeError("groupings %s were not computed by any rank" % missing)
    lens = [c - 1 for c in counts]
    offs = [0] * (n + 1)
    for i in range(n):
        offs[i + 1] = offs[i] + lens[i]
    flat = torch.zeros(max(offs[n], 1), dtype=torch.int64)
    for i, res in local.items():
        if lens[i]:
            flat[offs[i]:offs[i + 1]] = torch.as_tensor(list(res), dtype=torch.int64)
    dist.all_reduce(flat)
    out = flat.tolist()
    return [out[offs[i]:offs[i + 1]] for i in range(n)]


# ---- numpy's global RNG state as a token passed along the groupings ------------------------------
# The seed draws of grouping g continue the stream where grouping g-1 left it (the reference draws
# them one grouping after the other in the main process, set_cover_filter.py:824-827).  Instead of
# every rank replaying every grouping's draws, the state travels: the owner of grouping g receives
# it from the owner of g-1, draws, and sends it on before it starts the device work.
def _state_tensor():
    import numpy as np
    torch = sys.modules['torch']
    name, key, pos, has_gauss, cached = np.random.get_state()
    if name != 'MT19937':
        raise RuntimeError("numpy's legacy global RNG is expected to be MT19937")
    t = torch.empty(627, dtype=torch.int64)
    t[:624] = torch.from_numpy(key.astype(np.int64))
    t[624], t[625] = int(pos), int(has_gauss)
    t[626] = int(np.array([cached], dtype=np.float64).view(np.int64)[0])
    return t


def _set_state_from(t):
    import numpy as np
    a = t.numpy()
    cached = float(np.array([a[626]], dtype=np.int64).view(np.float64)[0])
    np.random.set_state(('MT19937', a[:624].astype(np.uint32), int(a[624]), int(a[625]), cached))


def rng_recv(src):
    """Block until the RNG state arrives from rank `src` and install it."""
    dist, torch = _dist(), sys.modules['torch']
    t = torch.empty(627, dtype=torch.int64)
    dist.recv(t, src=src)
    _set_state_from(t)


def rng_isend(dst):
    """Send the current RNG state to rank `dst` without waiting; returns (request, tensor): keep
    both alive and wait() on the request before the process group is torn down."""
    t = _state_tensor()
    return _dist().isend(t, dst=dst), t


def rng_broadcast(src):
    """Everyone ends with the state rank `src` holds."""
    dist, torch = _dist(), sys.modules['torch']
    t = _state_tensor() if dist.get_rank() == src else torch.empty(627, dtype=torch.int64)
    dist.broadcast(t, src=src)
    if dist.get_rank() != src:
        _set_state_from(t)


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


def ensure_exchange(ctx, need_bytes, token=0, failed=False):
    """Collective over all ranks, once per sharded set cover: make every rank's exchange area (the
    peer-mapped memory of cb_setcover_sharded) at least as large as the largest `need_bytes` of any
    rank, mapping all areas again when one had to grow.  The same small all-reduce carries two
    checks: `token` (a checksum of host state that must be identical on every rank, e.g. of numpy's
    RNG state the seed draws came from) and `failed` (this rank hit an error and will raise after
    the exchange), so that a rank-local problem turns into an exception on EVERY rank instead of
    leaving the others waiting in the kernel's barrier."""
    dist = _dist()
    if dist is None:
        raise RuntimeError("probe sharding needs an initialised torch.distributed process group")
    torch = sys.modules['torch']
    rank, world_size = dist.get_rank(), dist.get_world_size()
    token = int(token) & 0x3fffffffffffffff
    t = torch.tensor([int(need_bytes), token, -token, 1 if failed else 0], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # CPU tensor: gloo
    need, tmax, tmin, any_failed = int(t[0]), int(t[1]), -int(t[2]), int(t[3])
    if any_failed:
        raise RuntimeError("a rank failed before the sharded set cover" + (" (this one)" if failed else ""))
    if tmax != tmin:
        raise RuntimeError("ranks disagree on the host state a sharded grouping depends on: numpy's RNG state (seed "
                           "np.random identically on every rank before calling the filter) or the probe list itself "
                           "(a list ordered by a Python set differs between processes unless PYTHONHASHSEED is pinned)")
    if getattr(ctx, 'exchange_ready', False) and ctx.exchange_bytes() >= need and \
            getattr(ctx, 'exchange_n_ranks', 0) == world_size:
        return
    ctx.exchange_alloc(max(need + need // 2, 64 << 20))
    handle, _ = ctx.exchange_handle()
    handles = [None] * world_size
    dist.all_gather_object(handles, handle)
    ctx.exchange_attach(rank, world_size, handles=handles)
    # nobody launches a sharded kernel before every rank has zeroed and mapped its area
    dist.barrier()


def rng_state_token():
    """64-bit checksum of numpy's legacy global RNG state (see ensure_exchange)."""
    import zlib
    import numpy as np
    name, key, pos, has_gauss, cached = np.random.get_state()
    return (zlib.crc32(key.tobytes()) << 20) ^ (int(pos) << 1) ^ int(has_gauss)
