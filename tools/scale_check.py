#!/usr/bin/env python
"""Scale / sanity runs on the GPU (not a test): near-duplicate filter + set cover filter at a given
size, timings, and size-independent checks (the selection covers the whole universe; the two modes
of the greedy kernel agree)."""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from catch_b200 import _lib, probe  # noqa: E402
from catch_b200 import coverage as cov  # noqa: E402
from catch_b200.filter.near_duplicate_filter import NearDuplicateFilterWithMinHash  # noqa: E402
from catch_b200.filter.set_cover_filter import SetCoverFilter  # noqa: E402
from tests import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='influenza', choices=['influenza', 'zika'])
    ap.add_argument('--genomes', type=int, default=500)
    ap.add_argument('--ndf', type=float, default=0.6)
    ap.add_argument('--no-ndf', action='store_true')
    ap.add_argument('--check', action='store_true')
    a = ap.parse_args()
    ctx = _lib.default_context()
    t0 = time.perf_counter()
    if a.shape == 'influenza':
        gens = helpers.synthetic_influenza(a.genomes, seed=3)
        groups = [[[seg] for g in gens for seg in g]]          # every segment record is its own Genome
        pl, ps, scf_kw = 100, 50, dict(mismatches=5, lcf_thres=30, cover_extension=50)
    else:
        seqs = helpers.synthetic_genomes(a.genomes, 11000, 0.03, 2)
        groups = [[[s] for s in seqs]]
        pl, ps, scf_kw = 75, 50, dict(mismatches=2, lcf_thres=60, cover_extension=50)
    genomes = helpers.to_genomes(groups)
    cands = helpers.tile_candidates([s for g in groups[0] for s in g], pl, ps)
    probes = [[probe.Probe.from_str(s) for s in cands]]
    T = sum(len(s) for g in groups[0] for s in g)
    print('generated: %d genomes/universes, T=%d bp, P_raw=%d (%.1f s)' % (len(groups[0]), T, len(cands),
                                                                           time.perf_counter() - t0), flush=True)
    np.random.seed(7)
    random.seed(7)
    if not a.no_ndf:
        ndf = NearDuplicateFilterWithMinHash(a.ndf)
        t = time.perf_counter()
        probes = ndf.filter(probes, genomes, input_is_grouped=True)
        dt = time.perf_counter() - t
        st = ndf.last_stats
        print('NDF: %d -> %d probes in %.3f s (%.0f probes/s); device %.1f ms (signatures %.1f, rounds %.1f x%d), '
              'distance checks %d' % (len(cands), len(probes[0]), dt, len(cands) / dt, st['ms_total'],
                                      st['ms_seed_index'], st['ms_greedy'], st['n_picks'], st['n_candidate_hits']),
              flush=True)
    else:
        probes = [list(dict.fromkeys(probes[0]))]
    scf = SetCoverFilter(**scf_kw)
    t = time.perf_counter()
    out = scf.filter(probes, genomes, input_is_grouped=True)
    dt = time.perf_counter() - t
    s = scf.last_stats[0]
    P = len(probes[0])
    print('SCF: P=%d -> %d probes in %.3f s  (%.3g pairs/s e2e)' % (P, len(out[0]), dt, P * T / dt))
    print('  coverage %s' % {k: round(v, 2) if isinstance(v, float) else v for k, v in s['coverage'].items()
                              if k != 'reserved'})
    print('  setcover %s' % {k: round(v, 2) if isinstance(v, float) else v for k, v in s['setcover'].items()
                              if k != 'reserved'}, s['setcover']['reserved'][:4], flush=True)
    if a.check:
        strs = [p.seq_str for p in probes[0]]
        np.random.seed(7)
        random.seed(7)
        if not a.no_ndf:
            NearDuplicateFilterWithMinHash(a.ndf)._params()        # advance `random` like the first run
        group = cov.PackedGroup(ctx, strs, genomes[0])
        plan = cov.SeedPlan(strs, scf_kw['mismatches'], scf_kw['lcf_thres'], 20)
        cover, _ = cov.compute_cover(ctx, group, plan, scf_kw['mismatches'], scf_kw['lcf_thres'], 0,
                                     scf_kw['cover_extension'])
        picks_inc, _ = ctx.setcover(cover, len(strs))
        os.environ['CB_SETCOVER_FULL'] = '1'
        picks_full, _ = ctx.setcover(cover, len(strs))
        del os.environ['CB_SETCOVER_FULL']
        print('  incremental vs full greedy picks identical:', picks_inc.tolist() == picks_full.tolist())
        pid, gen, st_, en_ = ctx.cover_export(cover)
        sel = np.zeros(len(strs), dtype=bool)
        sel[picks_inc] = True
        n_g = len(genomes[0])
        glen = np.array([g.size() for g in genomes[0]], dtype=np.int64)
        base = np.concatenate(([0], np.cumsum(glen)))
        uni = np.zeros(base[-1] + 1, dtype=np.int32)
        cov_sel = np.zeros(base[-1] + 1, dtype=np.int32)
        np.add.at(uni, base[gen] + st_, 1)
        np.add.at(uni, base[gen] + en_, -1)
        m = sel[pid]
        np.add.at(cov_sel, base[gen[m]] + st_[m], 1)
        np.add.at(cov_sel, base[gen[m]] + en_[m], -1)
        u = np.cumsum(uni)[:-1] > 0
        c = np.cumsum(cov_sel)[:-1] > 0
        print('  universe bits %d, covered by selection %d, selection covers universe: %s' %
              (u.sum(), (u & c).sum(), bool(np.all(c[u]))))
        sel_strs = sorted(strs[i] for i in picks_inc.tolist())
        print('  filter output equals picks:', sorted(p.seq_str for p in out[0]) == sel_strs)
        cover.free()
        group.free()


if __name__ == '__main__':
    main()
