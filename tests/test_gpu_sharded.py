"""The probe-sharded multi-GPU path (cb_coverage_range + cb_setcover_sharded, csrc/rounds.cu).

On a one-GPU box the ranks are several contexts of this process sharing the device ("virtual
ranks": same kernels, same exchange protocol through each other's exchange areas, each persistent
grid capped so that all fit on the device together); with two or more GPUs the real one-process-
per-GPU path is run under torchrun as well.  Needs a B200."""
import os
import random
import subprocess
import sys
import threading

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(['nvidia-smi', '-L'], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith('GPU '))
    except Exception:
        return 0


def _workload(case):
    n, length, pl, kw, dedup = [
        (60, 4000, 75, dict(mismatches=2, lcf_thres=60, cover_extension=50), True),
        (40, 3000, 100, dict(mismatches=5, lcf_thres=30, cover_extension=0), True),
        (25, 2000, 75, dict(mismatches=0, lcf_thres=75, cover_extension=0), False),   # ties, duplicate probes
    ][case]
    seqs = helpers.synthetic_genomes(n, length, 0.03, 10 + case)
    cands = helpers.tile_candidates(seqs, pl, 50)
    if dedup:
        cands = list(dict.fromkeys(cands))
    return seqs, cands, kw


def _sharded_picks(ctxs, cands, seqs, kw, plan, grid_limit, ranks=None):
    """Run the sharded path with len(ctxs) virtual ranks (threads); returns per-rank (picks, stats)."""
    from catch_b200 import coverage as cov
    from catch_b200 import parallel
    R, P = len(ctxs), len(cands)
    out, err = [None] * R, [None] * R
    covers, groups = [None] * R, [None] * R
    try:
        need = 0
        for r, c in enumerate(ctxs):
            lo, hi = parallel.shard_bounds(P, R, r)
            groups[r] = cov.PackedGroup(c, cands, [[s] for s in seqs])
            covers[r], _ = cov.compute_cover_range(c, groups[r], plan, kw['mismatches'], kw['lcf_thres'], 0,
                                                   kw['cover_extension'], lo, hi)
            need = max(need, c.exchange_required(covers[r]))
        for c in ctxs:
            if c.exchange_bytes() < need:
                c.exchange_alloc(need)
        # the ranks share one device and one memory pool here: make sure no rank has to go to the driver for
        # memory (which waits for running kernels) while another rank's persistent kernel waits for it
        ctxs[0].pool_reserve(R * (need + (64 << 20)))
        addrs = [c.exchange_handle()[1] for c in ctxs]
        for r, c in enumerate(ctxs):
            c.exchange_attach(r, R, addresses=addrs, grid_limit=grid_limit)

        # all host-side set-up first, then the persistent kernels: with the ranks sharing ONE device, a driver
        # call of one rank's set-up (memory, module state) can otherwise wait for another rank's running kernel
        jobs = [None] * R
        for r, c in enumerate(ctxs):
            lo, hi = parallel.shard_bounds(P, R, r)
            jobs[r] = c.setcover_sharded_begin(covers[r], lo, hi, ranks)

        def run(r):
            try:
                out[r] = ctxs[r].setcover_sharded_end(jobs[r], P)
            except BaseException as e:      # noqa: BLE001 -- reported by the main thread
                err[r] = e
        threads = [threading.Thread(target=run, args=(r,)) for r in range(R)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        for x in covers + groups:
            if x is not None:
                x.free()
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize('n_ranks', [2, 4])
def test_sharded_setcover_virtual_ranks(ctx, n_ranks):
    """Stage A on shards of the probes + the sharded greedy loop give, on every rank, exactly the pick
    sequence of the one-GPU path (which is pinned against the oracle elsewhere)."""
    from catch_b200 import _lib
    from catch_b200 import coverage as cov
    ctxs = [_lib.Context(0) for _ in range(n_ranks)]
    try:
        for case in range(3):
            seqs, cands, kw = _workload(case)
            np.random.seed(7)
            random.seed(7)
            plan = cov.SeedPlan(cands, kw['mismatches'], kw['lcf_thres'], 20)
            group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
            cover, _ = cov.compute_cover(ctx, group, plan, kw['mismatches'], kw['lcf_thres'], 0, kw['cover_extension'])
            want, st1 = ctx.setcover(cover, len(cands))          # also loads every kernel the ranks will launch
            full = ctx.cover_export(cover)
            cover.free()
            group.free()
            assert len(want) > 0
            got = _sharded_picks(ctxs, cands, seqs, kw, plan, grid_limit=296 // n_ranks)
            for r in range(n_ranks):
                assert got[r][0].tolist() == want.tolist(), (case, r)
            # ranks agree on the round structure too (same candidate lists everywhere)
            assert len({(int(s.reserved[4]), int(s.reserved[5])) for _, s in got}) == 1
            del full
    finally:
        for c in ctxs:
            c.close()


def test_coverage_range_tiles_the_full_cover(ctx):
    """cb_coverage_range over the shards of a probe list = the rows of the full cover."""
    from catch_b200 import coverage as cov
    from catch_b200 import parallel
    seqs, cands, kw = _workload(0)
    np.random.seed(7)
    plan = cov.SeedPlan(cands, kw['mismatches'], kw['lcf_thres'], 20)
    group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
    cover, _ = cov.compute_cover(ctx, group, plan, kw['mismatches'], kw['lcf_thres'], 0, kw['cover_extension'])
    full = [a.tolist() for a in ctx.cover_export(cover)]
    cover.free()
    parts = [[], [], [], []]
    for r in range(3):
        lo, hi = parallel.shard_bounds(len(cands), 3, r)
        c, _ = cov.compute_cover_range(ctx, group, plan, kw['mismatches'], kw['lcf_thres'], 0, kw['cover_extension'],
                                       lo, hi)
        ex = ctx.cover_export(c)
        c.free()
        assert all(lo <= p < hi for p in ex[0].tolist())
        for k in range(4):
            parts[k] += ex[k].tolist()
    group.free()
    assert parts == full


def test_sharded_setcover_with_ranks_virtual(ctx):
    """Ranks (identify / avoided genomes) in the sharded loop: rank by rank, as on one GPU."""
    from catch_b200 import _lib
    from catch_b200 import coverage as cov
    seqs, cands, kw = _workload(0)
    rng = np.random.default_rng(5)
    ranks = rng.integers(0, 3, len(cands)).astype(np.int32)
    np.random.seed(7)
    plan = cov.SeedPlan(cands, kw['mismatches'], kw['lcf_thres'], 20)
    group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
    cover, _ = cov.compute_cover(ctx, group, plan, kw['mismatches'], kw['lcf_thres'], 0, kw['cover_extension'])
    want, _ = ctx.setcover(cover, len(cands), ranks=ranks)
    cover.free()
    group.free()
    ctxs = [_lib.Context(0) for _ in range(2)]
    try:
        got = _sharded_picks(ctxs, cands, seqs, kw, plan, grid_limit=148, ranks=ranks)
        for r in range(2):
            assert got[r][0].tolist() == want.tolist()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.skipif(_n_gpus() < 2, reason='needs two GPUs')
def test_probe_sharded_filter_two_processes():
    """One process per GPU under torchrun: SetCoverFilter in probe-sharded mode returns what the
    group-sharded / single-GPU path returns, on every rank (tools/multigpu_check.py)."""
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29541',
                          os.path.join(ROOT, 'tools', 'multigpu_check.py')],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and 'MULTIGPU_CHECK OK' in res.stdout, (res.stdout[-2000:], res.stderr[-2000:])
