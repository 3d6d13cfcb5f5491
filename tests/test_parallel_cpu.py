"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group, device work replaced by a
stub (no CUDA call is made).  Checks that groupings are partitioned, that every rank ends with the
complete result, and that the RNG stream a rank sees does not depend on which groupings it owns."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_groups_is_deterministic_and_balanced():
    from catch_b200 import parallel
    sizes = [5, 100, 7, 90, 3, 60, 60]
    own = parallel.assign_groups(sizes, 3)
    assert own == parallel.assign_groups(sizes, 3)
    load = [sum(s for s, o in zip(sizes, own) if o == r) for r in range(3)]
    assert sum(load) == sum(sizes) and max(load) <= 120 and min(load) >= 100
    assert parallel.assign_groups(sizes, 1) == [0] * len(sizes)
    assert parallel.assign_groups([], 4) == []


def _worker(rank, world_size, port, out_path, fail_group=None):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import random
    import torch.distributed as dist
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    from tests import helpers
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    computed = []

    def fake_select(self, group_i, n_groups, probe_strs, group, plan, plan_tol, target_genomes, all_genomes):
        if group_i == fail_group:
            raise ValueError("grouping %d cannot be designed" % group_i)
        computed.append(group_i)
        # deterministic stand-in for the device result: depends on the drawn seeds
        return [int(x) % max(1, len(probe_strs)) for x in plan.seed_pos[:3]]

    SetCoverFilter._select_for_group = fake_select

    # no device in this test: stand-ins for the packed group and the context
    from catch_b200 import coverage as cov

    class FakeGroup:
        def __init__(self, ctx, probe_strs, genomes, gathered=None, targets_staged=None):
            self.probe_len = np.array([len(s) for s in probe_strs], dtype=np.int32)
            self.probes = None

        def free(self):
            pass

    class FakeCtx:
        def probes_have_duplicates(self, probes):
            return True

    cov.PackedGroup = FakeGroup
    SetCoverFilter._context = lambda self: FakeCtx()
    rng = random.Random(5)
    groups = [helpers.random_groups(rng, n_groups=1)[0] for _ in range(5)]
    cands = [helpers.tile_candidates([s for g in gens for s in g], 30, 10) for gens in groups]
    probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
    f = SetCoverFilter(mismatches=1, lcf_thres=20)          # random seed mode: consumes np.random
    np.random.seed(3)
    if fail_group is not None:
        try:
            f.filter(probes, helpers.to_genomes(groups), input_is_grouped=True)
            outcome = 'no exception'
        except Exception as e:
            outcome = type(e).__name__ + ': ' + str(e)
        with open(out_path, 'w') as fh:
            fh.write(outcome)
        dist.destroy_process_group()
        return
    out = f.filter(probes, helpers.to_genomes(groups), input_is_grouped=True)
    after = int(np.random.randint(0, 1 << 30))
    np.save(out_path, np.array([computed, [[p.seq_str for p in g] for g in out], after], dtype=object),
            allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('world_size,draw', [(1, 'replay'), (2, 'replay'), (2, 'chain')])
def test_group_sharding_over_gloo(tmp_path, monkeypatch, world_size, draw):
    """draw: how a group-sharded run gets its seed draws -- every rank replays the whole stream (default) or the
    RNG state travels from owner to owner (CB_DRAW=chain); both must reproduce the single-process stream."""
    import multiprocessing as mp
    monkeypatch.setenv('CB_DRAW', draw)
    ctxm = mp.get_context('spawn')
    port = _free_port()
    paths = [str(tmp_path / ('r%d.npy' % r)) for r in range(world_size)]
    procs = [ctxm.Process(target=_worker, args=(r, world_size, port, paths[r])) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = [np.load(p, allow_pickle=True) for p in paths]
    all_computed = sorted(g for r in res for g in r[0])
    assert all_computed == [0, 1, 2, 3, 4]                    # every grouping exactly once
    for r in res[1:]:
        assert r[1] == res[0][1]                              # every rank has the full result
        assert r[2] == res[0][2]                              # and saw the same RNG stream
    # world size must not change the result or the stream either
    ref_path = tmp_path / 'single.npy'
    if world_size > 1:
        p = ctxm.Process(target=_worker, args=(0, 1, _free_port(), str(ref_path)))
        p.start()
        p.join(180)
        assert p.exitcode == 0
        single = np.load(str(ref_path), allow_pickle=True)
        assert single[1] == res[0][1] and single[2] == res[0][2]


def test_failure_on_one_rank_raises_on_every_rank(tmp_path):
    """A grouping that fails on its owner must not leave the other ranks waiting in a collective:
    the failure travels with the results and every rank raises."""
    import multiprocessing as mp
    ctxm = mp.get_context('spawn')
    port = _free_port()
    paths = [str(tmp_path / ('f%d.txt' % r)) for r in range(2)]
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, paths[r], 2)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    outcomes = [open(p).read() for p in paths]
    assert sum(o.startswith('ValueError: grouping 2') for o in outcomes) == 1
    assert sum(o.startswith('RuntimeError: rank') for o in outcomes) == 1


# ---- probe-sharded path: the host-side collective that precedes cb_setcover_sharded ----------------
def _exchange_worker(rank, world, port, out_path, scenario):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from catch_b200 import parallel

    class FakeCtx:
        """Stands in for _lib.Context: records what the collective asks of the exchange area."""

        def __init__(self):
            self.bytes, self.exchange_ready, self.log = 0, False, []

        def exchange_bytes(self):
            return self.bytes

        def exchange_alloc(self, n):
            self.bytes = n
            self.log.append(('alloc', n))

        def exchange_handle(self):
            return bytes([rank]) * 64, 0

        def exchange_attach(self, r, n, handles=None, addresses=None, grid_limit=0):
            self.log.append(('attach', r, n, [h[0] for h in handles]))
            self.exchange_ready, self.exchange_rank, self.exchange_n_ranks = True, r, n

    ctx = FakeCtx()
    res = []
    try:
        if scenario == 'grow':
            # sizes differ per rank: everyone must end up with the largest; a second, smaller request keeps the area
            parallel.ensure_exchange(ctx, (100 << 20) * (rank + 1), token=5)
            parallel.ensure_exchange(ctx, 1 << 20, token=6)
            parallel.ensure_exchange(ctx, (400 << 20) + rank, token=7)
            res = ctx.log
        elif scenario == 'token':
            parallel.ensure_exchange(ctx, 1 << 20, token=11 + rank)        # ranks disagree on host state
        elif scenario == 'failed':
            parallel.ensure_exchange(ctx, 1 << 20, token=3, failed=(rank == 1))
        elif scenario == 'lists_agree':
            # group-sharded result exchange: both ranks hold the same lists (same fingerprint)
            res = parallel.exchange_group_results({rank: [3 + rank, 1]}, [0, 1], rank, None, token=77)
        elif scenario == 'lists_differ':
            # ... or lists in a different order (a set-ordered list without a pinned PYTHONHASHSEED)
            parallel.exchange_group_results({rank: [3 + rank, 1]}, [0, 1], rank, None, token=77 + rank)
        outcome = 'ok'
    except Exception as e:
        outcome = type(e).__name__ + ': ' + str(e)
    np.save(out_path, np.array([outcome, res], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('scenario', ['grow', 'token', 'failed', 'lists_agree', 'lists_differ'])
def test_ensure_exchange_over_gloo(tmp_path, scenario):
    """World size 2 over gloo: the exchange areas grow to the largest need of any rank and are mapped
    again only then; a rank-local failure or diverging RNG state raises on EVERY rank."""
    import multiprocessing as mp
    ctxm = mp.get_context('spawn')
    port = _free_port()
    paths = [str(tmp_path / ('x%d.npy' % r)) for r in range(2)]
    procs = [ctxm.Process(target=_exchange_worker, args=(r, 2, port, paths[r], scenario)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = [np.load(p, allow_pickle=True) for p in paths]
    if scenario == 'grow':
        for r, (outcome, log) in enumerate(res):
            assert outcome == 'ok'
            allocs = [e[1] for e in log if e[0] == 'alloc']
            attaches = [e for e in log if e[0] == 'attach']
            assert len(allocs) == 2 and len(attaches) == 2          # the 1 MiB request reused the area
            assert allocs[0] >= (200 << 20) and allocs[1] >= (400 << 20) + 1
            assert attaches[0][1:] == (r, 2, [0, 1])
        assert [e[1] for e in res[0][1] if e[0] == 'alloc'] == [e[1] for e in res[1][1] if e[0] == 'alloc']
    elif scenario == 'token':
        assert all(o[0].startswith('RuntimeError: ranks disagree') for o in res)
    elif scenario == 'lists_agree':
        assert all(o[0] == 'ok' and o[1] == [[3, 1], [4, 1]] for o in res)
    elif scenario == 'lists_differ':
        assert all(o[0].startswith('RuntimeError: ranks hold different probe lists') for o in res)
    else:
        assert all(o[0].startswith('RuntimeError: a rank failed') for o in res)
        assert '(this one)' in res[1][0] and '(this one)' not in res[0][0]
