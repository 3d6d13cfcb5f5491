#!/usr/bin/env python
"""BASELINE config 4 (V-All shape) through SetCoverFilter: many independent taxa = many groupings,
sharded over the GPUs of one box (one process per GPU, no data-path collective; the selected ids are
exchanged once).  Run single-process for the 1-GPU number, under torchrun for N GPUs:

    python tools/vall_bench.py --taxa 32
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/vall_bench.py --taxa 32

The full config has 300 taxa x 333 genomes; --taxa scales it down (state the size with the result).
Prints one JSON line (rank 0): pairs/s = sum_g P_g * T_g / wall time of SetCoverFilter.filter()
(max over ranks), -pl 100 -m 5 -l 30 -e 0 as in SURVEY.md 8(d)."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--taxa', type=int, default=32)
    ap.add_argument('--genomes', type=int, default=333)
    ap.add_argument('--reps', type=int, default=2)
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('cpu:gloo,cuda:nccl')
    from catch_b200 import _lib, probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    os.environ['CB_SHARD'] = 'groups'
    t0 = time.perf_counter()
    groups = helpers.synthetic_taxa(a.taxa, a.genomes, seed=4)
    genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
    cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
    probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
    pairs = sum(len(c) * sum(map(len, g)) for c, g in zip(cands, groups))
    t_gen = time.perf_counter() - t0
    ctx = _lib.Context(local)
    scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=0)
    scf._ctx = ctx
    best, n_sel = None, None
    for rep in range(a.reps + 1):                      # first repetition is the warm-up
        np.random.seed(7)
        random.seed(7)
        if dist is not None:
            dist.barrier(device_ids=[local])
        t = time.perf_counter()
        out = scf.filter(probes, genomes, input_is_grouped=True)
        dt = time.perf_counter() - t
        if dist is not None:
            import torch
            tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if rep > 0 and (best is None or dt < best):
            best = dt
        n_sel = sum(len(o) for o in out)
    dev_ms = sum((s['coverage']['ms_total'] + s['setcover']['ms_total']) for s in scf.last_stats if s and 'coverage' in s)
    if rank == 0:
        print(json.dumps({
            'workload': 'config 4 shape (V-All): %d taxa x %d genomes of 10-30 kb, -pl 100 -m 5 -l 30 -e 0' % (a.taxa, a.genomes),
            'n_gpus': world, 'groupings': a.taxa, 'P_total': sum(map(len, cands)),
            'T_total_bp': sum(sum(map(len, g)) for g in groups), 'pairs': pairs, 'selected': n_sel,
            'e2e_s': best, 'pairs_per_s_e2e': pairs / best, 'rank0_device_ms': round(dev_ms, 1),
            'sharding': 'groupings over ranks, largest first (catch_b200/parallel.py); no data-path collective',
            'generation_s': round(t_gen, 1)}), flush=True)
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
