#!/usr/bin/env python
"""More oracle goldens at BASELINE shapes (see make_scale_golden.py for the first two):

  config4_6x40 : V-All shape, 6 taxa x 40 genomes of 10-30 kb (helpers.synthetic_taxa, generator seed 4), one
                 grouping per taxon, -pl 100 -ps 50 -m 5 -l 30 -e 0, numpy seed 7 ONCE before the call: the seed draws
                 of grouping g continue the stream where grouping g-1 stopped (set_cover_filter.py:824-827), which is
                 what the device path has to reproduce when it works on several groupings at a time;
  config5_40   : the hybridisation sweep on the config-3 input of make_scale_golden.py (40 influenza-shaped genomes
                 after the MinHash near-duplicate filter): m in {0, 2, 10} x l in {100, 60, 30} cells, numpy seed 7
                 per cell.

Generated with the CPU oracle; a few minutes on 8 threads.

    PYTHONHASHSEED=0 python tests/golden/make_scale_golden2.py
"""
import gzip
import hashlib
import json
import os
import random
import sys

if os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from tests import helpers  # noqa: E402


def md5(strs):
    return hashlib.md5('\n'.join(strs).encode()).hexdigest()


def main():
    threads = max(1, os.cpu_count() or 1)
    out = {}
    groups = helpers.synthetic_taxa(6, 40, seed=4)
    cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
    np.random.seed(7)
    random.seed(7)
    sel = O.set_cover_filter(cands, [[[s] for s in g] for g in groups], 5, 30, 0, 1.0, 0, 20, n_threads=threads)
    out['config4_6x40'] = dict(n_taxa=6, n_genomes=40, gen_seed=4, pl=100, ps=50,
                               scf=dict(mismatches=5, lcf_thres=30, cover_extension=0), np_seed=7,
                               n_cands=[len(c) for c in cands], cands_md5=[md5(c) for c in cands],
                               selected=[[int(x) for x in s] for s in sel],
                               rng_after=int(np.random.randint(0, 1 << 30)))
    print('config4_6x40:', [len(c) for c in cands], '->', [len(s) for s in sel], flush=True)
    with gzip.open(os.path.join(HERE, 'scale_oracle.json.gz'), 'rt') as f:
        c3 = json.load(f)['config3_40']
    gens = helpers.synthetic_influenza(c3['n_genomes'], seed=c3['gen_seed'])
    segs = [seg for g in gens for seg in g]
    tiles = helpers.tile_candidates(segs, c3['pl'], c3['ps'])
    scf_in = [tiles[i] for i in c3['kept_idx']]
    cells = []
    for m in (0, 2, 10):
        for l in (100, 60, 30):
            np.random.seed(7)
            s5, det = O.set_cover_filter([scf_in], [[[s] for s in segs]], m, l, 0, 1.0, 50, 20, n_threads=threads,
                                         return_details=True)
            cells.append(dict(m=m, l=l, n_intervals=int(len(det[0]['quads'])), picks=[int(x) for x in det[0]['picks']],
                              selected=[int(x) for x in s5[0]]))
            print('config5_40 m=%d l=%d: %d picks' % (m, l, len(s5[0])), flush=True)
    out['config5_40'] = dict(cells=cells, np_seed=7, cover_extension=50)
    with gzip.open(os.path.join(HERE, 'scale_oracle2.json.gz'), 'wt') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
