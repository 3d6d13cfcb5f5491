"""Live comparison of the oracle with the reference imported from /root/reference (skipped where
that tree is absent, e.g. on the GPU box).  CPU only."""
import os
import random
import sys

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'catch')), reason='reference tree not present')

from tests import helpers  # noqa: E402


@pytest.fixture(scope='module')
def ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import catch.probe
    import catch.genome
    import catch.filter.set_cover_filter
    import catch.utils.longest_common_substring
    import catch.utils.set_cover
    import catch.utils.interval
    return sys.modules['catch']


def test_k_lcf_random(ref):
    from oracle import oracle as O
    rng = random.Random(1)
    for _ in range(400):
        n = rng.randint(5, 60)
        a = ''.join(rng.choice('ACGTN') for _ in range(n))
        b = ''.join((rng.choice('ACGTN') if rng.random() < 0.15 else c) for c in a)
        a2, b2 = a + 'A' * rng.randint(0, 5), b + 'C' * rng.randint(0, 5)
        s = rng.randint(0, n - 2)
        e = rng.randint(s + 1, min(n, s + 8))
        b2 = b2[:s] + a2[s:e] + b2[e:]
        k = rng.randint(0, 4)
        want = ref.utils.longest_common_substring.k_lcf_around_anchor(a2, b2, s, e, k)
        assert O.k_lcf_around_anchor(a2, b2, s, e, k) == (int(want[0]), int(want[1]))


def test_choose_seeds_consumes_rng_like_reference(ref):
    """The batched randint replay in catch_b200.probe equals the reference's per-probe
    np.random.choice calls (probe.py:393-396), including mixed probe lengths."""
    from catch_b200 import probe as bprobe
    rng = random.Random(3)
    for trial in range(5):
        lens = [rng.choice([75, 75, 75, 60, 100]) if trial % 2 else 75 for _ in range(200)]
        strs = [''.join(rng.choice('ACGT') for _ in range(L)) for L in lens]
        probes = [ref.probe.Probe.from_str(s) for s in strs]
        np.random.seed(trial)
        m = ref.probe._construct_rand_kmer_probe_map(probes, k=20, include_positions=True)
        want = {}
        for kmer, hits in m.items():
            for p, pos in hits:
                want.setdefault(p.seq_str, set()).add(pos)
        np.random.seed(trial)
        k, seeds, mode = bprobe.choose_seed_positions(lens, 2, 30)
        assert mode == 'random' and k == 20
        got = {}
        for s, row in zip(strs, seeds):
            got.setdefault(s, set()).update(int(x) for x in row)
        assert got == want


def test_set_cover_filter_random_vs_reference(ref):
    from collections import OrderedDict
    from oracle import oracle as O
    for case in range(200, 206):
        groups, cands, params = helpers.random_case(case)
        refg = [[ref.genome.Genome.from_one_seq(s[0]) if len(s) == 1 else
                 ref.genome.Genome.from_chrs(OrderedDict((str(i), x) for i, x in enumerate(s))) for s in gens]
                for gens in groups]
        probes = [[ref.probe.Probe.from_str(s) for s in c] for c in cands]
        f = ref.filter.set_cover_filter.SetCoverFilter(**params)
        f._force_num_processes = 1
        np.random.seed(case)
        random.seed(case)
        out = f.filter(probes, refg, input_is_grouped=True)
        want = []
        for gi, go in zip(probes, out):
            idmap = {id(p): i for i, p in enumerate(gi)}
            want.append([idmap[id(p)] for p in go])
        np.random.seed(case)
        random.seed(case)
        got = O.set_cover_filter(cands, groups, params['mismatches'], params['lcf_thres'],
                                 params['island_of_exact_match'], params['coverage'],
                                 params['cover_extension'], params['kmer_probe_map_k'])
        assert got == want
