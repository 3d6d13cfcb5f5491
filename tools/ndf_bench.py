#!/usr/bin/env python
"""Near-duplicate filter + set cover filter on the config-3 (influenza-shaped) input, repeated, with the
host/device split of each call (GPU, not a test).

    python tools/ndf_bench.py [--genomes 5000] [--reps 3]"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from catch_b200 import _lib, probe  # noqa: E402
from catch_b200.filter.near_duplicate_filter import NearDuplicateFilterWithMinHash  # noqa: E402
from catch_b200.filter.set_cover_filter import SetCoverFilter  # noqa: E402
from tests import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--genomes', type=int, default=5000)
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    ctx = _lib.default_context()
    gens = helpers.synthetic_influenza(a.genomes, seed=3)
    groups = [[[seg] for g in gens for seg in g]]
    genomes = helpers.to_genomes(groups)
    cands = helpers.tile_candidates([s for g in groups[0] for s in g], 100, 50)
    T = sum(len(s) for g in groups[0] for s in g)
    raw = [[probe.Probe.from_str(s) for s in cands]]
    print('config 3 shape: %d genomes x 8 segments, T=%d bp, P_raw=%d' % (a.genomes, T, len(cands)), flush=True)
    for rep in range(a.reps):
        np.random.seed(7)
        random.seed(7)
        ndf = NearDuplicateFilterWithMinHash(0.6)
        ndf._ctx = ctx
        t = time.perf_counter()
        kept = ndf.filter(raw, genomes, input_is_grouped=True)
        dt = time.perf_counter() - t
        st = ndf.last_stats
        print('rep %d NDF: %d -> %d distinct -> %d kept in %.1f ms (%.2e probes/s); device %.1f ms = grouping %.1f + '
              'signatures %.1f + %d rounds %.1f' % (rep, len(cands), st['n_distinct'], len(kept[0]), dt * 1e3,
                                                   len(cands) / dt, st['ms_total'], st['ms_pack'], st['ms_seed_index'],
                                                   st['n_picks'], st['ms_greedy']), flush=True)
        scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=50)
        scf._ctx = ctx
        t = time.perf_counter()
        out = scf.filter(kept, genomes, input_is_grouped=True)
        dt = time.perf_counter() - t
        s = scf.last_stats[0]
        P = len(kept[0])
        print('rep %d SCF: %d -> %d in %.1f ms (%.2e pairs/s e2e); device %.1f ms; host %s' % (
            rep, P, len(out[0]), dt * 1e3, P * T / dt, s['coverage']['ms_total'] + s['setcover']['ms_total'],
            {k: round(v, 1) for k, v in s['host_ms'].items()}), flush=True)


if __name__ == '__main__':
    main()
