import sys, os, time, random
sys.path.insert(0, os.getcwd())
import numpy as np
from tests import helpers
from catch_b200 import _lib, probe
from catch_b200.filter.set_cover_filter import SetCoverFilter
from catch_b200.probe_batch import ProbeBatch
n_taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 6
groups = helpers.synthetic_taxa(n_taxa, 333, seed=4)
genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
batches = [ProbeBatch(np.frombuffer(''.join(c).encode(), dtype=np.uint8).reshape(len(c), 100)) for c in cands]
ctx = _lib.default_context()
if os.environ.get("RESERVE"): ctx.pool_reserve(int(os.environ["RESERVE"]) << 30)
for pf in ('lists', 'batch', 'lists', 'batch'):
    scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=0); scf._ctx = ctx
    np.random.seed(7); random.seed(7)
    t = time.perf_counter(); out = scf.filter(probes if pf == 'lists' else batches, genomes, input_is_grouped=True); dt = time.perf_counter() - t
    print(pf, round(dt * 1e3, 1))
    for s in scf.last_stats:
        c = s['coverage']
        print('   P', s['n_probes'], 'mode', s['seed_mode'], 'raw', c['n_raw_ranges'], 'iv', c['n_intervals'], 'scan', round(c['ms_scan_emit'], 2),
              'merge', round(c['ms_merge'], 2), 'launches', c['n_kernel_launches'], 'host', {k: round(v, 1) for k, v in s['host_ms'].items()}, flush=True)
