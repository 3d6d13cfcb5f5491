"""genome clustering by MinHash sketches on the device, against fixtures recorded by RUNNING
THE REFERENCE (tests/golden/make_f3_golden.py: catch/utils/cluster.py + catch/utils/lsh.py) and against the
oracle on seeded random inputs.  Bit-exact: signatures (integers), distances (the Python doubles), the float32
condensed matrix, the clusters and their order.  Needs a B200."""
import random

import numpy as np
import pytest

from catch_b200 import _lib
from catch_b200.utils import cluster
from oracle import oracle
from tests import golden_io, helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold():
    return golden_io.load('f3_reference.json.gz')


def test_sketches_match_reference(ctx, gold):
    for c in gold['sketches']:
        h = cluster.SketchFunction(c['k'], c['N'], c['a'], c['b'])
        sk = h.sketch(c['seqs'].values(), ctx)
        assert sk.signatures().tolist() == c['sigs'], c['name']
        # one sequence at a time gives the same tuples (make_h's h(s))
        first = next(iter(c['seqs'].values()))
        assert list(h(first)) == c['sigs'][0]


def test_distances_match_reference(ctx, gold):
    for c in gold['sketches']:
        n = len(c['sigs'])
        sk = cluster.SketchSet.from_signatures(np.array(c['sigs'], dtype=np.uint32), ctx)
        d = sk.rows(list(range(n)))
        want = np.array(c['dist'], dtype=np.float64)
        assert d.shape == want.shape
        assert np.array_equal(d, want), c['name']                       # the same doubles
        assert np.array_equal(sk.condensed(), np.array(c['condensed'], dtype=np.float32)), c['name']
        assert sk(0, n - 1) == want[0, n - 1]
        fam = cluster.MinHashFamily(c['k'], N=c['N'])
        assert fam.estimate_jaccard_dist(tuple(c['sigs'][0]), tuple(c['sigs'][1])) == want[0, 1]


def test_clusters_match_reference(ctx, gold):
    for c in gold['clusters']:
        random.seed(c['seed'])
        got = cluster.cluster_with_minhash_signatures(c['seqs'], k=c['k'], N=c['N'], threshold=c['threshold'],
                                                      cluster_method=c['method'])
        assert got == c['clusters'], (c['name'], c['method'], c['threshold'])


def test_make_signatures_consumes_random_like_reference(ctx, gold):
    c = gold['sketches'][0]
    random.seed(3)
    sigs = cluster.make_signatures_with_minhash(cluster.MinHashFamily(c['k'], N=c['N']), c['seqs'])
    assert [list(sigs[h]) for h in c['seqs']] == c['sigs']


@pytest.mark.parametrize('seed,k,N', [(1, 12, 100), (2, 12, 1), (3, 5, 64), (4, 31, 500), (5, 55, 1024), (6, 12, 100)])
def test_sketches_match_oracle_on_random_inputs(ctx, seed, k, N):
    rng = random.Random(seed)
    seqs = []
    for fam in range(3):
        anc = ''.join(rng.choice('ACGT') for _ in range(rng.randint(200, 1500)))
        for _ in range(rng.randint(1, 5)):
            s = helpers.mutate(rng, anc, rng.choice([0.0, 0.01, 0.1]), 'ACGTN' if seed % 2 else 'ACGT')
            seqs.append(s[:rng.randint(max(k, 60), len(s))])
    seqs.append(''.join(rng.choice('ACGT') for _ in range(k)))            # exactly one k-mer
    seqs.append('A' * (k + 40))                                           # one distinct k-mer, 41 times
    seqs.append(''.join(rng.choice('AC') for _ in range(k + N // 2)))     # fewer k-mers than N
    a, b = rng.randint(1, 2 ** 31 - 1), rng.randint(0, 2 ** 31 - 1)
    if seed == 6:
        a, b = 2 ** 31 - 1, 2 ** 31 - 1                                   # both drawable (randint is inclusive)
    sk = cluster.SketchFunction(k, N, a, b).sketch(seqs, ctx)
    got = sk.signatures()
    want = [list(oracle.sketch(s, k, N, a, b)) for s in seqs]
    assert got.tolist() == want
    d = sk.rows(list(range(len(seqs))))
    for i in range(len(seqs)):
        for j in range(len(seqs)):
            assert d[i, j] == oracle.sketch_jaccard_dist(want[i], want[j], N)


def test_many_sequences_against_oracle(ctx):
    """Several hundred sequences: batches, a CTA per sequence, both clustering methods."""
    seqs = {}
    for f in range(6):
        for i, s in enumerate(helpers.synthetic_genomes(40, 900 + 150 * f, 0.02 + 0.01 * f, 100 + f)):
            seqs['f%d_%d' % (f, i)] = s
    for method in ('simple', 'hierarchical'):
        random.seed(5)
        got = cluster.cluster_with_minhash_signatures(seqs, threshold=0.12, cluster_method=method)
        random.seed(5)
        want = oracle.cluster_with_minhash_signatures(seqs, threshold=0.12, cluster_method=method)
        assert got == want


def test_sequence_shorter_than_kmer_is_rejected(ctx):
    with pytest.raises(AssertionError):
        cluster.SketchFunction(12, 100, 5, 7).sketch(['ACGTACGTACG'], ctx)
    with pytest.raises(_lib.CatchB200Error):
        ctx.sketch_sequences(b'ACGTACGTACG', np.array([0, 11], dtype=np.int64), 12, 100, 5, 7)
    with pytest.raises(_lib.CatchB200Error):
        ctx.sketch_sequences(b'ACGTACGTACGTT', np.array([0, 13], dtype=np.int64), 12, 2000, 5, 7)


def test_near_rows_are_the_thresholded_distance_rows(ctx, gold):
    """cb_sketch_near_rows: the part of cb_sketch_dist_rows within the threshold, same doubles, ascending columns."""
    for c in gold['sketches']:
        n = len(c['sigs'])
        sk = cluster.SketchSet.from_signatures(np.array(c['sigs'], dtype=np.uint32), ctx)
        full = sk.rows(list(range(n)))
        for thr in (0.0, 0.3, 0.8, 1.0):
            rows = list(range(n - 1, -1, -1))
            off, idx, dist = ctx.sketch_near_rows(sk.h, rows, thr)
            assert off[0] == 0 and len(off) == n + 1 and off[-1] == len(idx) == len(dist)
            for t, r in enumerate(rows):
                want = np.flatnonzero(full[r] <= thr)
                assert idx[off[t]:off[t + 1]].tolist() == want.tolist(), (c['name'], thr, r)
                assert np.array_equal(dist[off[t]:off[t + 1]], full[r][want])
