"""cProfile of SetCoverFilter.filter() on the V-All shape (16 taxa), lists of Probe objects and ProbeBatch input:
where the host time of a many-groupings run goes.  python tools/vall_profile.py [lists|batch]"""
import cProfile
import io
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402

from catch_b200 import _lib, probe  # noqa: E402
from catch_b200.filter.set_cover_filter import SetCoverFilter  # noqa: E402
from catch_b200.probe_batch import ProbeBatch  # noqa: E402
from tests import helpers  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'lists'
groups = helpers.synthetic_taxa(16, 333, seed=4)
genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
if mode == 'lists':
    inp = [[probe.Probe.from_str(s) for s in c] for c in cands]
else:
    inp = [ProbeBatch(np.frombuffer(''.join(c).encode(), dtype=np.uint8).reshape(len(c), 100)) for c in cands]
ctx = _lib.default_context()
scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=0)
scf._ctx = ctx
for rep in range(3):
    np.random.seed(7)
    random.seed(7)
    pr = cProfile.Profile() if rep == 2 else None
    t = time.perf_counter()
    if pr:
        pr.enable()
    out = scf.filter(inp, genomes, input_is_grouped=True)
    if pr:
        pr.disable()
    print(mode, 'rep', rep, round((time.perf_counter() - t) * 1e3, 1), 'ms', sum(len(o) for o in out), 'selected', flush=True)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22)
print(s.getvalue()[:6000])
