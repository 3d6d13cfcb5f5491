import sys, os, time, random, json
sys.path.insert(0, os.getcwd())
import numpy as np
from tests import helpers
from catch_b200 import _lib, probe
from catch_b200.filter.set_cover_filter import SetCoverFilter
groups = helpers.synthetic_taxa(16, 333, seed=4)
genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
ctx = _lib.default_context()
from catch_b200.probe_batch import ProbeBatch
batches = [ProbeBatch(np.frombuffer(''.join(c).encode(), dtype=np.uint8).reshape(len(c), 100)) for c in cands]
for pf in ('lists', 'batch', 'lists', 'batch'):
    scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=0); scf._ctx = ctx
    np.random.seed(7); random.seed(7)
    t = time.perf_counter(); out = scf.filter(probes if pf == 'lists' else batches, genomes, input_is_grouped=True); dt = time.perf_counter() - t
    agg = {}
    for s in scf.last_stats:
        for k in ('ms_seed_index', 'ms_scan_count', 'ms_scan_emit', 'ms_merge', 'ms_total'):
            agg['cov_' + k] = agg.get('cov_' + k, 0) + s['coverage'][k]
        for k in ('ms_universe', 'ms_greedy', 'ms_total'):
            agg['sc_' + k] = agg.get('sc_' + k, 0) + s['setcover'][k]
        for k, v in s.get('host_ms', {}).items():
            agg['host_' + k] = agg.get('host_' + k, 0) + v
    print(pf, round(dt * 1e3, 1), {k: round(v, 1) for k, v in agg.items()}, flush=True)
