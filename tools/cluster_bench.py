#!/usr/bin/env python
"""Genome clustering by MinHash sketches at a stated size: sketch kernels, distance rows, the whole
cluster_with_minhash_signatures call.  Prints one JSON line.  Also the command the ncu capture of the sketch kernels
is taken from.

    python tools/cluster_bench.py [--taxa 4] [--genomes 333] [--method simple|hierarchical]
"""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from catch_b200 import _lib  # noqa: E402
from catch_b200.utils import cluster  # noqa: E402
from tests import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--taxa', type=int, default=4)
    ap.add_argument('--genomes', type=int, default=333)
    ap.add_argument('--method', default='simple')
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    groups = helpers.synthetic_taxa(args.taxa, args.genomes, seed=4)
    seqs = {'t%d_%d' % (t, i): s for t, g in enumerate(groups) for i, s in enumerate(g)}
    bases = sum(map(len, seqs.values()))
    ctx = _lib.default_context()
    best = None
    for _ in range(args.reps):
        random.seed(7)
        t = time.perf_counter()
        clusters = cluster.cluster_with_minhash_signatures(seqs, threshold=0.15, cluster_method=args.method)
        dt = time.perf_counter() - t
        st = cluster.cluster_with_minhash_signatures.last_stats
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    h = cluster.SketchFunction(12, 100, 12345, 678).sketch(seqs.values(), ctx)
    t = time.perf_counter()
    rows = h.rows(list(range(min(len(seqs), 256))))
    t_rows = time.perf_counter() - t
    t = time.perf_counter()
    cond = h.condensed()
    t_cond = time.perf_counter() - t
    print(json.dumps({
        'sequences': len(seqs), 'bases': bases, 'method': args.method, 'clusters': [len(c) for c in clusters],
        'wall_ms': dt * 1e3, 'hash_kernel_ms': st['ms_scan_emit'], 'select_kernel_ms': st['ms_merge'],
        'device_ms': st['ms_total'], 'gbases_per_s_hash_kernel': bases / st['ms_scan_emit'] / 1e6,
        'rows_256_ms': t_rows * 1e3, 'condensed_ms': t_cond * 1e3, 'pairs': int(cond.size)}))


if __name__ == '__main__':
    main()
