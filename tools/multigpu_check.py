#!/usr/bin/env python
"""Run under torchrun with >= 2 GPUs: the probe-sharded mode (stage A split over ranks, coverage
all-gathered over NCCL) must select exactly what the group-sharded / single-GPU path selects.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py
"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from catch_b200 import _lib, probe  # noqa: E402
from catch_b200.filter.set_cover_filter import SetCoverFilter  # noqa: E402
from tests import helpers  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('cpu:gloo,cuda:nccl')
    ctx = _lib.Context(local)
    ok = True
    for case, (n, length, pl, kw) in enumerate([
            (60, 4000, 75, dict(mismatches=2, lcf_thres=60, cover_extension=50)),
            (40, 3000, 100, dict(mismatches=5, lcf_thres=30, cover_extension=0)),
            (25, 2000, 75, dict(mismatches=0, lcf_thres=75)),
            (3, 500, 75, dict(mismatches=1, lcf_thres=50, coverage=0.7))]):
        seqs = helpers.synthetic_genomes(n, length, 0.03, 10 + case)
        cands = helpers.tile_candidates(seqs, pl, 50)
        if case % 2 == 0:
            cands = list(dict.fromkeys(cands))       # odd cases keep duplicate probes
        genomes = helpers.to_genomes([[[s] for s in seqs]])
        probes = [[probe.Probe.from_str(s) for s in cands]]
        out = {}
        for mode in ('groups', 'probes'):
            os.environ['CB_SHARD'] = mode
            f = SetCoverFilter(**kw)
            f._ctx = ctx
            np.random.seed(7)
            random.seed(7)
            t = time.perf_counter()
            res = f.filter(probes, genomes, input_is_grouped=True)
            dt = time.perf_counter() - t
            out[mode] = [p.seq_str for p in res[0]]
            if rank == 0:
                print('case %d mode %-6s: %d probes -> %d selected in %.1f ms' % (case, mode, len(cands), len(res[0]), dt * 1e3))
        same = out['groups'] == out['probes']
        gathered = [None] * world
        dist.all_gather_object(gathered, out['probes'])
        same = same and all(g == gathered[0] for g in gathered)
        ok = ok and same
        if rank == 0:
            print('case %d identical across modes and ranks: %s' % (case, same))
    if rank == 0:
        print('MULTIGPU_CHECK', 'OK' if ok else 'FAILED')
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
