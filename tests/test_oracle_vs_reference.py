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


def test_tiling_fasta_and_cluster_oracle_against_the_live_reference(ref, tmp_path):
    """The callers in front of the filters, live: vectorised candidate tiling, the native FASTA
    parser, and the oracle's sketch / clustering restatement against the reference's own modules."""
    import catch.filter.candidate_probes as rcp
    import catch.utils.cluster as rcluster
    import catch.utils.seq_io as rseq_io
    from catch_b200.probe_batch import ProbeBatch
    from catch_b200.utils import seq_io
    from oracle import oracle
    import logging
    logging.disable(logging.WARNING)
    try:
        rng = random.Random(31)
        for _ in range(60):
            seqs = []
            for _ in range(rng.randint(1, 8)):
                n = rng.randint(60, 400)
                s = [rng.choice('ACGT') for _ in range(n)]
                if rng.random() < 0.4:
                    for _ in range(rng.randint(1, 3)):
                        a = rng.randrange(n)
                        for i in range(a, min(n, a + rng.choice([1, 2, 3, 10]))):
                            s[i] = 'N'
                seqs.append(''.join(s))
            L, st = rng.choice([40, 60]), rng.choice([7, 20, 30, 60])
            want = rcp.make_candidate_probes_from_sequences(seqs, probe_length=L, probe_stride=st)
            got = ProbeBatch.from_sequences(seqs, L, st)
            assert [(p.seq_str, p.is_flanking_n_string) for p in got] == [(p.seq_str, p.is_flanking_n_string) for p in want]
        fn = str(tmp_path / 'x.fasta')
        with open(fn, 'w') as f:
            f.write('>a b\r\nacgt-ry\r\nNNtt\n\n>c\nAC GT\t\n>a b\nTTTT\n')
        assert list(seq_io.read_fasta(fn).items()) == list(rseq_io.read_fasta(fn).items())
        seqs = {}
        for g in range(3):
            for i, s in enumerate(helpers.synthetic_genomes(5, 1200, 0.05, 50 + g)):
                seqs['g%d_%d' % (g, i)] = s
        for method in ('simple', 'hierarchical'):
            random.seed(2)
            want = rcluster.cluster_with_minhash_signatures(seqs, threshold=0.12, cluster_method=method)
            random.seed(2)
            assert oracle.cluster_with_minhash_signatures(seqs, threshold=0.12, cluster_method=method) == want
    finally:
        logging.disable(logging.NOTSET)
