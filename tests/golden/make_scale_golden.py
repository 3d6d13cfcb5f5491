#!/usr/bin/env python
"""Pick-sequence goldens at BASELINE shapes too large for the per-test oracle run:

  config2_120 : 120 genomes x 11 kb, div 0.03, generator seed 2 (the first 120 genomes of config 2's
                generator), -pl 75 -ps 50 -m 2 -l 60 -e 50, numpy seed 7;
  config3_40  : 40 influenza-shaped genomes x 8 segments (generator seed 3), pl 100 ps 50, MinHash
                near-duplicate filter 0.6 under random.seed(7), then -m 5 -l 30 -e 50, numpy seed 7.

Generated with the CPU oracle (oracle/, itself pinned against the reference by
tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py); about 1.5 minutes on 8 threads.

    PYTHONHASHSEED=0 python tests/golden/make_scale_golden.py

The fixture stores generator parameters, not sequences: tests regenerate the inputs with
tests/helpers.py and check their checksum first.
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
    # ---- config 2 shape, 120 genomes
    seqs = helpers.synthetic_genomes(120, 11000, 0.03, 2)
    cands = list(dict.fromkeys(helpers.tile_candidates(seqs, 75, 50)))
    np.random.seed(7)
    random.seed(7)
    sel, det = O.set_cover_filter([cands], [[[s] for s in seqs]], 2, 60, 0, 1.0, 50, 20, n_threads=threads,
                                  return_details=True)
    out['config2_120'] = dict(n_genomes=120, length=11000, div=0.03, gen_seed=2, pl=75, ps=50,
                              scf=dict(mismatches=2, lcf_thres=60, cover_extension=50), np_seed=7,
                              n_cands=len(cands), cands_md5=md5(cands), n_intervals=int(len(det[0]['quads'])),
                              picks=[int(x) for x in det[0]['picks']], selected=[int(x) for x in sel[0]])
    print('config2_120: %d candidates, %d picks' % (len(cands), len(sel[0])), flush=True)
    # ---- config 3 shape, 40 genomes, near-duplicate filter first
    gens = helpers.synthetic_influenza(40, seed=3)
    segs = [seg for g in gens for seg in g]
    c3 = helpers.tile_candidates(segs, 100, 50)
    random.seed(7)
    kept = O.near_duplicate_minhash(c3, 0.6)
    first = {}
    for i, s in enumerate(c3):
        first.setdefault(s, i)
    kept_idx = sorted(first[s] for s in kept)
    scf_in = [c3[i] for i in kept_idx]
    np.random.seed(7)
    sel3, det3 = O.set_cover_filter([scf_in], [[[s] for s in segs]], 5, 30, 0, 1.0, 50, 20, n_threads=threads,
                                    return_details=True)
    out['config3_40'] = dict(n_genomes=40, gen_seed=3, pl=100, ps=50, ndf=dict(dist_thres=0.6, random_seed=7),
                             scf=dict(mismatches=5, lcf_thres=30, cover_extension=50), np_seed=7,
                             n_cands=len(c3), cands_md5=md5(c3), kept_idx=kept_idx,
                             n_intervals=int(len(det3[0]['quads'])),
                             picks=[int(x) for x in det3[0]['picks']], selected=[int(x) for x in sel3[0]])
    print('config3_40: %d candidates -> %d kept -> %d picks' % (len(c3), len(kept_idx), len(sel3[0])), flush=True)
    with gzip.open(os.path.join(HERE, 'scale_oracle.json.gz'), 'wt') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
