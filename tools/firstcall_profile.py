import sys, time, cProfile, pstats, random
sys.path.insert(0, '/root/repo')
import numpy as np
from catch_b200 import _lib, probe
from catch_b200.filter.set_cover_filter import SetCoverFilter
from tests import helpers
seqs = helpers.synthetic_genomes(100, 5000, 0.03, 2)
genomes = helpers.to_genomes([[[s] for s in seqs]])
cands = list(dict.fromkeys(helpers.tile_candidates(seqs, 75, 50)))
probes = [[probe.Probe.from_str(s) for s in cands]]
t=time.perf_counter(); ctx=_lib.default_context(); print('ctx init %.3f s' % (time.perf_counter()-t))
f = SetCoverFilter(2, 60, cover_extension=50)
for i in range(3):
    np.random.seed(7)
    pr = cProfile.Profile()
    t=time.perf_counter()
    pr.enable(); out = f.filter(probes, genomes, input_is_grouped=True); pr.disable()
    print('call %d: %.3f s' % (i, time.perf_counter()-t))
    if i == 0:
        pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
