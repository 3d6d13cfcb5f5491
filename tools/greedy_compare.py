#!/usr/bin/env python
"""Stage B kernels side by side on one cover (GPU, not a test): pick sequences must be identical;
prints time, rounds and phase timers of each kernel.

    python tools/greedy_compare.py [--genomes 500] [--m 2 --lcf 60 --pl 75 --e 50] [--caps 256,2048]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from catch_b200 import _lib  # noqa: E402
from catch_b200 import coverage as cov  # noqa: E402
from tests import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--genomes', type=int, default=500)
    ap.add_argument('--length', type=int, default=11000)
    ap.add_argument('--m', type=int, default=2)
    ap.add_argument('--lcf', type=int, default=60)
    ap.add_argument('--pl', type=int, default=75)
    ap.add_argument('--e', type=int, default=50)
    ap.add_argument('--caps', default='2048')
    ap.add_argument('--modes', default='inc,par')
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    ctx = _lib.default_context()
    seqs = helpers.synthetic_genomes(a.genomes, a.length, 0.03, 2)
    cands = list(dict.fromkeys(helpers.tile_candidates(seqs, a.pl, 50)))
    group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
    np.random.seed(7)
    plan = cov.SeedPlan(cands, a.m, a.lcf, 20)
    cover, st_a = cov.compute_cover(ctx, group, plan, a.m, a.lcf, 0, a.e)
    print('P=%d T=%d intervals=%d scan %.2f ms' % (len(cands), sum(map(len, seqs)), st_a.n_intervals,
                                                   st_a.ms_scan_emit), flush=True)
    ref = None
    runs = [(m, None) for m in a.modes.split(',') if m != 'par']
    if 'par' in a.modes.split(','):
        runs += [('par', c) for c in a.caps.split(',')]
    for mode, cap in runs:
        os.environ['CB_GREEDY'] = mode
        if cap:
            os.environ['CB_GREEDY_LIST_CAP'] = cap
        best = None
        for _ in range(a.reps):
            ctx.flush_l2()
            picks, st = ctx.setcover(cover, len(cands), None, None)
            if best is None or st.ms_greedy < best.ms_greedy:
                best = st
        r = list(best.reserved)
        same = None if ref is None else (picks.tolist() == ref)
        if ref is None:
            ref = picks.tolist()
        print('%-4s cap=%-5s picks=%d greedy %.3f ms (universe/index %.3f ms)  phases ms: rebuild %.2f mark+check %.2f '
              'apply %.2f barrier %.2f  rebuilds=%d rounds=%d active/round=%.0f  same_as_first=%s' % (
                  mode, cap, len(picks), best.ms_greedy, best.ms_universe, r[0] / 1e6, r[1] / 1e6, r[2] / 1e6, r[3] / 1e6,
                  r[4], r[5], (r[6] / r[5]) if r[5] else 0, same), flush=True)
    cover.free()
    group.free()


if __name__ == '__main__':
    main()
