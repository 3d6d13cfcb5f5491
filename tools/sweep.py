#!/usr/bin/env python
"""BASELINE config 5: hybridisation sweep m in {0,2,5,10} x l in {30,60,100} on the config-3
(influenza-shaped) input, pl 100, MinHash near-duplicate filter first.  Prints one JSON line per
cell: pairs/s through SetCoverFilter (e2e and device-only), seed mode, kernel times, and the scan
kernel's operand traffic rate against the measured HBM peak.

  python tools/sweep.py --genomes 1000          (the full config uses 5000 genomes)
"""
import argparse
import json
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
    ap.add_argument('--genomes', type=int, default=1000)
    ap.add_argument('--repeat', type=int, default=3)
    a = ap.parse_args()
    ctx = _lib.default_context()
    gens = helpers.synthetic_influenza(a.genomes, seed=3)
    groups = [[[seg] for g in gens for seg in g]]
    genomes = helpers.to_genomes(groups)
    cands = helpers.tile_candidates([s for g in groups[0] for s in g], 100, 50)
    T = sum(len(s) for g in groups[0] for s in g)
    np.random.seed(7)
    random.seed(7)
    ndf = NearDuplicateFilterWithMinHash(0.6)
    t = time.perf_counter()
    probes = ndf.filter([[probe.Probe.from_str(s) for s in cands]], genomes, input_is_grouped=True)
    t_ndf = time.perf_counter() - t
    P = len(probes[0])
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        peak = 6650.0
    print(json.dumps({'input': 'config 3 shape, %d genomes x 8 segments' % a.genomes, 'T_bp': T,
                      'P_raw': len(cands), 'P_after_minhash_ndf': P, 'ndf_s': round(t_ndf, 3)}), flush=True)
    for m in (0, 2, 5, 10):
        for l in (30, 60, 100):
            scf = SetCoverFilter(mismatches=m, lcf_thres=l, cover_extension=50)
            scf._ctx = ctx
            best = None
            for _ in range(a.repeat):
                np.random.seed(7)
                ctx.flush_l2()
                t = time.perf_counter()
                out = scf.filter(probes, genomes, input_is_grouped=True)
                dt = time.perf_counter() - t
                if best is None or dt < best[0]:
                    best = (dt, dict(scf.last_stats[0]), len(out[0]))
            dt, s, n_sel = best
            ca, cb = s['coverage'], s['setcover']
            dev_ms = ca['ms_total'] + cb['ms_total']
            nw, planes = 2, s['bits']
            scan_ms = ca['ms_scan_emit']
            operand = 2 * T * planes / 8 + ca['n_candidate_hits'] * (8 + (planes + 1) * nw * 8) + ca['n_raw_ranges'] * 16
            print(json.dumps({
                'm': m, 'l': l, 'seed_mode': s['seed_mode'], 'k': s['k'], 'selected': n_sel,
                'pairs_per_s_e2e': P * T / dt, 'pairs_per_s_device': P * T / (dev_ms / 1e3),
                'e2e_ms': round(dt * 1e3, 2), 'device_ms': round(dev_ms, 2), 'scan_ms': round(scan_ms, 2),
                'merge_ms': round(ca['ms_merge'], 2), 'greedy_ms': round(cb['ms_greedy'], 2),
                'hits': ca['n_candidate_hits'], 'intervals': ca['n_intervals'], 'picks': cb['n_picks'],
                'greedy_rounds': cb['reserved'][5], 'greedy_rebuilds': cb['reserved'][4],
                'scan_operand_GBps': round(operand / (scan_ms / 1e3) / 1e9, 1) if scan_ms > 0 else None,
                'scan_operand_frac_of_hbm_peak': round(operand / (scan_ms / 1e3) / 1e9 / peak, 3) if scan_ms > 0 else None,
            }), flush=True)


if __name__ == '__main__':
    main()
