#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Two sources, both executed from /root/reference (never copied):
  1. the reference's own unit tests for the hot path are run with thin recording wrappers
     around the hot-path entry points; every (input, output) pair the tests exercise is
     captured, so the fixtures hold the golden vectors of
       catch/utils/tests/test_longest_common_substring.py, catch/utils/tests/test_set_cover.py,
       catch/tests/test_probe.py, catch/filter/tests/test_set_cover_filter.py,
       catch/filter/tests/test_near_duplicate_filter.py
     as data (inputs and the reference's outputs), not as code;
  2. seeded random cases and the BASELINE config-1 design are pushed through the reference.

Run:  PYTHONHASHSEED=0 python tests/golden/make_golden.py      (needs /root/reference)
The fixtures are consumed by tests/test_oracle_golden.py (CPU, oracle) and
tests/test_gpu_golden.py (B200, CUDA path).
"""
import gzip
import hashlib
import json
import os
import random
import sys
import unittest

if os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from catch import probe as rprobe  # noqa: E402
from catch import genome as rgenome  # noqa: E402
from catch.filter import near_duplicate_filter as rndf  # noqa: E402
from catch.filter import set_cover_filter as rscf  # noqa: E402
from catch.filter import duplicate_filter as rdf  # noqa: E402
from catch.filter import candidate_probes as rcp  # noqa: E402
from catch.filter import probe_designer as rpd  # noqa: E402
from catch.utils import interval as rinterval  # noqa: E402
from catch.utils import longest_common_substring as rlcs  # noqa: E402
from catch.utils import set_cover as rsc  # noqa: E402

from tests import helpers  # noqa: E402

REC = {'lcs': [], 'lcf': [], 'scan': [], 'setcover': [], 'scf': [], 'ndf': []}
CAP = {'lcs': 400, 'lcf': 300, 'scan': 150, 'setcover': 200, 'scf': 200, 'ndf': 60}
SEEN = {k: set() for k in REC}


def _add(kind, rec):
    key = json.dumps(rec, sort_keys=True)
    if key in SEEN[kind] or len(REC[kind]) >= CAP[kind]:
        return
    SEEN[kind].add(key)
    REC[kind].append(rec)


def _s(x):
    return x if isinstance(x, str) else ''.join(x)


def _pairs(x):
    return [[int(a), int(b)] for a, b in x]


# ----------------------------------------------------------------------------- wrappers
_orig_lcs = rlcs.k_lcf_around_anchor
_direct = {'lcs': True}


def lcs_wrapper(a, b, anchor_start, anchor_end, k):
    out = _orig_lcs(a, b, anchor_start, anchor_end, k)
    if _direct['lcs']:
        _add('lcs', dict(a=_s(a), b=_s(b), s=int(anchor_start), e=int(anchor_end), k=int(k),
                         out=[int(out[0]), int(out[1])]))
    return out


_orig_factory = rprobe.probe_covers_sequence_by_longest_common_substring
_FN_PARAMS = {}


def factory_wrapper(mismatches, lcf_thres, island_of_exact_match=0):
    fn = _orig_factory(mismatches, lcf_thres, island_of_exact_match)

    def lcf(probe_seq, sequence, kmer_start, kmer_end, full_probe_len, full_sequence_len):
        prev = _direct['lcs']
        _direct['lcs'] = False
        try:
            out = fn(probe_seq, sequence, kmer_start, kmer_end, full_probe_len, full_sequence_len)
        finally:
            _direct['lcs'] = prev
        if prev:       # called by a test directly, not from inside a scan
            _add('lcf', dict(m=mismatches, lcf=lcf_thres, island=island_of_exact_match, p=_s(probe_seq),
                             s=_s(sequence), ks=int(kmer_start), ke=int(kmer_end), fpl=int(full_probe_len),
                             fsl=int(full_sequence_len), out=None if out is None else [int(out[0]), int(out[1])]))
        return out
    _FN_PARAMS[id(lcf)] = (mismatches, lcf_thres, island_of_exact_match, lcf)
    return lcf


_orig_open = rprobe.open_probe_finding_pool
_orig_find = rprobe.find_probe_covers_in_sequence
_POOL = {}


def open_wrapper(kmer_probe_map, fn, num_processes=None, use_native_dict=False):
    _POOL.clear()
    if id(fn) in _FN_PARAMS:
        m, l, isl, _ = _FN_PARAMS[id(fn)]
        seeds = {}
        for kmer, hits in kmer_probe_map.native_dict.items():
            for pstr, pos in hits:
                seeds.setdefault(pstr, set()).add(int(pos))
        _POOL.update(m=m, lcf=l, island=isl, k=int(kmer_probe_map.k),
                     probes=sorted(seeds), seeds=seeds)
    return _orig_open(kmer_probe_map, fn, num_processes, use_native_dict)


def find_wrapper(sequence, merge_overlapping=True):
    prev = _direct['lcs']
    _direct['lcs'] = False
    try:
        out = _orig_find(sequence, merge_overlapping)
    finally:
        _direct['lcs'] = prev
    if _POOL and len(sequence) <= 3000:
        _add('scan', dict(m=_POOL['m'], lcf=_POOL['lcf'], island=_POOL['island'], k=_POOL['k'],
                          probes=_POOL['probes'], seeds=[sorted(_POOL['seeds'][p]) for p in _POOL['probes']],
                          seq=sequence, merge=bool(merge_overlapping),
                          out={p.seq_str: _pairs(r) for p, r in out.items()}))
    return out


_orig_amu = rsc.approx_multiuniverse


def _intervals_of(v, use_arrays, use_intervalsets):
    if use_intervalsets:
        if isinstance(v, tuple):
            return [[int(v[0]), int(v[1])]]
        return _pairs(v.intervals)
    elems = sorted(set(int(x) for x in v))
    runs = []
    for x in elems:
        if runs and runs[-1][1] == x:
            runs[-1][1] = x + 1
        else:
            runs.append([x, x + 1])
    return runs


def amu_wrapper(sets, costs=None, universe_p=None, ranks=None, use_arrays=False, use_intervalsets=False,
                logger_prefix=""):
    out = _orig_amu(sets, costs=costs, universe_p=universe_p, ranks=ranks, use_arrays=use_arrays,
                    use_intervalsets=use_intervalsets, logger_prefix=logger_prefix)
    try:
        if all(isinstance(s, int) for s in sets) and sum(len(v) for v in sets.values()) < 4000:
            ukeys = []
            for v in sets.values():
                for u in v:
                    if u not in ukeys:
                        ukeys.append(u)
            uid = {u: i for i, u in enumerate(ukeys)}
            jsets = {str(s): {str(uid[u]): _intervals_of(x, use_arrays, use_intervalsets) for u, x in v.items()}
                     for s, v in sets.items()}
            _add('setcover', dict(
                sets=jsets,
                costs=None if costs is None else {str(s): float(c) for s, c in costs.items()},
                universe_p=None if universe_p is None else {str(uid[u]): float(p) for u, p in universe_p.items()
                                                             if u in uid},
                ranks=None if ranks is None else {str(s): int(r) for s, r in ranks.items()},
                out=sorted(int(x) for x in out)))
    except Exception as ex:       # never let recording break the reference run
        print('setcover record skipped:', ex)
    return out


_orig_scf_init = rscf.SetCoverFilter.__init__
_orig_scf_filter = rscf.SetCoverFilter._filter
_COUNTER = {'n': 0}


def scf_init_wrapper(self, *a, **kw):
    import inspect
    bound = inspect.signature(_orig_scf_init).bind(self, *a, **kw)
    bound.apply_defaults()
    self._rec_args = {k: v for k, v in bound.arguments.items() if k != 'self'}
    return _orig_scf_init(self, *a, **kw)


def scf_filter_wrapper(self, input, target_genomes_grouped):
    _COUNTER['n'] += 1
    seed = 1000 + _COUNTER['n']
    np.random.seed(seed)
    random.seed(seed)
    prev = _direct['lcs']
    _direct['lcs'] = False
    try:
        out = _orig_scf_filter(self, input, target_genomes_grouped)
    finally:
        _direct['lcs'] = prev
    a = self._rec_args
    simple = (a['custom_cover_range_fn'] is None and a['custom_cover_range_tolerant_fn'] is None)
    if simple:
        avoided = []
        for path in a['avoided_genomes']:
            from catch.utils import seq_io
            avoided.append(list(seq_io.iterate_fasta(path)))
        args = {k: v for k, v in a.items() if k not in ('avoided_genomes', 'custom_cover_range_fn',
                                                        'custom_cover_range_tolerant_fn')}
        ids = []
        for grp_in, grp_out in zip(input, out):
            idmap = {id(p): i for i, p in enumerate(grp_in)}
            ids.append([idmap[id(p)] for p in grp_out])
        _add('scf', dict(args=args, avoided=avoided, seed=seed,
                         probes=[[p.seq_str for p in g] for g in input],
                         genomes=[[list(g.seqs) for g in grp] for grp in target_genomes_grouped],
                         out=ids))
    return out


_orig_ndf_filter = rndf.NearDuplicateFilter._filter


def ndf_filter_wrapper(self, input):
    _COUNTER['n'] += 1
    seed = 5000 + _COUNTER['n']
    random.seed(seed)
    input = list(input)
    out = _orig_ndf_filter(self, input)
    if len(input) <= 3000:
        kind = 'minhash' if isinstance(self, rndf.NearDuplicateFilterWithMinHash) else 'hamming'
        _add('ndf', dict(kind=kind, dist_thres=self.dist_thres, k=self.k, reporting_prob=self.reporting_prob,
                         kmer_size=getattr(self.lsh_family, 'kmer_size', None),
                         dim=getattr(self.lsh_family, 'dim', None), seed=seed,
                         probes=[p.seq_str for p in input], out=[p.seq_str for p in out]))
    return out


def install():
    rlcs.k_lcf_around_anchor = lcs_wrapper
    rprobe.probe_covers_sequence_by_longest_common_substring = factory_wrapper
    rprobe.open_probe_finding_pool = open_wrapper
    rprobe.find_probe_covers_in_sequence = find_wrapper
    rsc.approx_multiuniverse = amu_wrapper
    rscf.SetCoverFilter.__init__ = scf_init_wrapper
    rscf.SetCoverFilter._filter = scf_filter_wrapper
    rndf.NearDuplicateFilter._filter = ndf_filter_wrapper
    # subclasses call NearDuplicateFilter._filter(self, input) explicitly, so the patch is seen


def run_reference_tests():
    names = ['catch.utils.tests.test_longest_common_substring', 'catch.utils.tests.test_set_cover',
             'catch.filter.tests.test_set_cover_filter', 'catch.filter.tests.test_near_duplicate_filter',
             'catch.tests.test_probe']
    suite = unittest.TestSuite()
    for n in names:
        suite.addTests(unittest.defaultTestLoader.loadTestsFromName(n))
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    print('reference tests: run=%d failures=%d errors=%d' % (res.testsRun, len(res.failures), len(res.errors)))
    return res


# ----------------------------------------------------------------------------- random cases
def random_cases():
    """Seeded random inputs through the (unwrapped) reference."""
    out = {'scf': [], 'ndf': [], 'hash': []}
    from collections import OrderedDict
    for case in range(16):
        groups, cands, params = helpers.random_case(case, alphabet='ACGT' if case % 4 else 'ACGTRYKMSW')
        refg = [[rgenome.Genome.from_one_seq(s[0]) if len(s) == 1 else
                 rgenome.Genome.from_chrs(OrderedDict((str(i), x) for i, x in enumerate(s))) for s in gens]
                for gens in groups]
        probes = [[rprobe.Probe.from_str(s) for s in c] for c in cands]
        f = rscf.SetCoverFilter(**params)
        f._force_num_processes = 1
        seed = 300 + case
        np.random.seed(seed)
        random.seed(seed)
        res = _orig_scf_filter(f, probes, refg)
        ids = []
        for gi, go in zip(probes, res):
            idmap = {id(p): i for i, p in enumerate(gi)}
            ids.append([idmap[id(p)] for p in go])
        out['scf'].append(dict(args=params, seed=seed, probes=cands, genomes=groups, out=ids))
    rng = random.Random(99)
    for case in range(10):
        L = rng.choice([50, 75, 100])
        bases = [''.join(rng.choice('ACGT') for _ in range(L)) for _ in range(rng.randint(3, 10))]
        ps = []
        for b in bases:
            for _ in range(rng.randint(1, 12)):
                ps.append(helpers.mutate(rng, b, rng.choice([0, 0.01, 0.03, 0.1])))
        ps += [rng.choice(ps) for _ in range(8)]
        rng.shuffle(ps)
        ref_in = [rprobe.Probe.from_str(s) for s in ps]
        seed = 700 + case
        if case % 2 == 0:
            d = rng.choice([0.3, 0.5, 0.6, 0.8])
            f = rndf.NearDuplicateFilterWithMinHash(d)
            random.seed(seed)
            res = _orig_ndf_filter(f, ref_in)
            out['ndf'].append(dict(kind='minhash', dist_thres=d, k=3, reporting_prob=0.8, kmer_size=10, dim=None,
                                   seed=seed, probes=ps, out=[p.seq_str for p in res]))
        else:
            d = rng.choice([0, 2, 5, 10])
            f = rndf.NearDuplicateFilterWithHammingDistance(d, L)
            random.seed(seed)
            res = _orig_ndf_filter(f, ref_in)
            out['ndf'].append(dict(kind='hamming', dist_thres=d, k=20, reporting_prob=0.8, kmer_size=None, dim=L,
                                   seed=seed, probes=ps, out=[p.seq_str for p in res]))
    for _ in range(200):
        s = ''.join(rng.choice('ACGTN') for _ in range(rng.randint(1, 40)))
        out['hash'].append([s, abs(hash(s))])
    return out


def config1_fingerprint():
    """BASELINE config 1 through the reference pipeline (candidates -> DuplicateFilter ->
    SetCoverFilter), as design.py wires it (bin/design.py:345-385)."""
    seqs = helpers.synthetic_genomes(20, 5000, 0.03, seed=1)
    genomes = [[rgenome.Genome.from_one_seq(s) for s in seqs]]
    filters = [rdf.DuplicateFilter(), rscf.SetCoverFilter(mismatches=0, lcf_thres=75, cover_extension=0)]
    pd = rpd.ProbeDesigner(genomes, filters, probe_length=75, probe_stride=50)
    np.random.seed(7)
    random.seed(7)
    pd.design()
    final = [p.seq_str for p in pd.final_probes]
    return dict(n_candidates=len(pd.candidate_probes), n_final=len(final),
                md5_sorted=hashlib.md5('\n'.join(sorted(final)).encode()).hexdigest(),
                md5_ordered=hashlib.md5('\n'.join(final).encode()).hexdigest())


def dump(name, obj):
    path = os.path.join(HERE, name)
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(json.dumps(obj, sort_keys=True).encode())
    print('%s: %d bytes' % (name, os.path.getsize(path)))


def main():
    rnd = random_cases()          # before the wrappers go in
    cfg1 = config1_fingerprint()
    install()
    run_reference_tests()
    for k, v in REC.items():
        print('recorded %s: %d' % (k, len(v)))
    dump('reference_tests.json.gz', REC)
    dump('random_cases.json.gz', rnd)
    dump('config1.json.gz', cfg1)


if __name__ == '__main__':
    main()
