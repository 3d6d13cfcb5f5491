"""Host-side logic that needs no GPU: C-ABI exports, candidate tiling, seed rule, FASTA I/O, CLI parser."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """libcatchb200.so loads (no CUDA call) and exports every function include/catch_b200.h declares."""
    from catch_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'catch_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(cb_[a-z0-9_]+)\s*\(', hdr)) - {'cb_status'})
    assert len(declared) >= 16
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert b'sm_100a' in _lib.load().cb_version()


def test_no_oracle_on_product_path():
    """Nothing under catch_b200/ or bin/ may import the oracle."""
    for base in ('catch_b200', 'bin'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    text = open(os.path.join(dirpath, f)).read()
                    assert 'oracle' not in text.lower(), os.path.join(dirpath, f)


def test_pigeonhole_rule():
    """probe.py:473-491 (pinned by the reference's tests/test_probe.py:270-335)."""
    from catch_b200 import probe
    assert probe.pigeonhole_kmer_length(100, 0, 20) == 100
    assert probe.pigeonhole_kmer_length(100, 2, 20) == 25      # k < 50, divides 100
    assert probe.pigeonhole_kmer_length(100, 3, 20) == 25
    assert probe.pigeonhole_kmer_length(75, 2, 20) == 25
    with pytest.raises(probe.PigeonholeRequiresTooSmallKmerSizeError):
        probe.pigeonhole_kmer_length(100, 5, 20)              # would need k < 20
    k, seeds, mode = probe.choose_seed_positions([75] * 3, 0, 75)
    assert (k, mode) == (75, 'pigeonhole') and seeds.tolist() == [[0]] * 3
    k, seeds, mode = probe.choose_seed_positions([100] * 2, 2, 100)
    assert (k, mode) == (25, 'pigeonhole') and seeds.tolist() == [[0, 25, 50, 75]] * 2
    np.random.seed(0)
    k, seeds, mode = probe.choose_seed_positions([75, 75, 60], 2, 60)
    assert (k, mode) == (20, 'random') and seeds.shape == (3, 20)
    assert seeds[:2].max() <= 55 and seeds[2].max() <= 40


def test_candidate_tiling_and_n_handling():
    from catch_b200.filter import candidate_probes as cp
    seq = 'ACGT' * 10 + 'A'            # 41 nt
    ps = [p.seq_str for p in cp.make_candidate_probes_from_sequence(seq, 10, 5)]
    assert ps[0] == seq[:10] and ps[-1] == seq[-10:] and len(ps) == 7 + 1
    with_n = 'ACGTACGTAC' + 'NN' + 'GTGTGTGTGT' + 'CCCC'
    got = [p.seq_str for p in cp.make_candidate_probes_from_sequence(with_n, 10, 4)]
    assert all('NN' not in s for s in got)
    assert 'ACGTACGTAC' in got and 'GTGTGTGTGT' in got       # flanking probes
    with pytest.raises(ValueError):
        cp.make_candidate_probes_from_sequence('ACGT', 10, 5)
    assert [p.seq_str for p in cp.make_candidate_probes_from_sequence('ACGTAC', 10, 5, allow_small_seqs=5)] == ['ACGTAC']


def test_fasta_round_trip(tmp_path):
    from catch_b200 import probe
    from catch_b200.utils import seq_io
    fn = tmp_path / 'in.fasta'
    fn.write_text('>a desc\nacgtRy-\nNNac\n\n>b\nTTTT\n')
    m = seq_io.read_fasta(str(fn))
    assert list(m.items()) == [('a desc', 'ACGTNNNNAC'), ('b', 'TTTT')]
    gs = seq_io.read_genomes_from_fasta(str(fn))
    assert [g.seqs for g in gs] == [['ACGTNNNNAC'], ['TTTT']] and gs[0].size() == 10
    out = tmp_path / 'out.fasta'
    p = probe.Probe.from_str('ACGTACGT')
    seq_io.write_probe_fasta([p], str(out))
    import hashlib
    assert out.read_text() == '>probe_%s\nACGTACGT\n' % hashlib.sha224(b'ACGTACGT').hexdigest()[-10:]


def test_cli_parser_matches_reference_defaults():
    sys.path.insert(0, os.path.join(ROOT, 'bin'))
    import design
    a = design.init_and_parse_args('basic', ['x.fasta', '-o', 'out.fasta'])
    assert (a.probe_length, a.probe_stride, a.mismatches, a.cover_extension, a.coverage) == (100, 50, 0, 0, 1.0)
    assert a.filter_with_lsh_minhash is None and a.lcf_thres is None
    b = design.init_and_parse_args('large', ['x.fasta', 'y.fasta', '-o', 'o.fa', '-pl', '75', '-c', '300'])
    assert (b.mismatches, b.cover_extension, b.filter_with_lsh_minhash, b.probe_length, b.coverage) == \
        (5, 50, 0.6, 75, 300)
    with pytest.raises(SystemExit):
        design.init_and_parse_args('basic', ['x.fasta', '-o', 'o', '-c', '1.5'])
    # genome clustering: off in design.py, on in design_large.py (bin/design.py:753-813 of the reference)
    assert (a.cluster_and_design_separately, a.cluster_from_fragments, a.cluster_and_design_separately_method) == \
        (None, None, 'choose')
    assert (b.cluster_and_design_separately, b.cluster_from_fragments, b.cluster_and_design_separately_method) == \
        (0.15, 50000, 'choose')
    with pytest.raises(SystemExit):
        design.init_and_parse_args('basic', ['x.fasta', '-o', 'o', '--cluster-and-design-separately', '0.6'])


def test_probe_designer_cluster_path_with_stub_filters(monkeypatch):
    """ProbeDesigner.design() with cluster_threshold (probe_designer.py:291-315): clusters become the groupings of the
    filters up to cluster_merge_after, the rest run ungrouped on the merged probes.  The clustering itself is stubbed
    (it needs the device); fragments, the 'choose' rule and the skip rule are the host logic under test."""
    from catch_b200 import genome
    from catch_b200.filter import probe_designer
    from catch_b200.filter.base_filter import BaseFilter
    from catch_b200.utils import cluster
    calls = {}

    def fake_cluster(seqs, threshold=None, cluster_method=None, **kw):
        calls['n'] = len(seqs)
        calls['method'] = cluster_method
        keys = list(seqs)
        return [keys[0::2], keys[1::2]]
    monkeypatch.setattr(cluster, 'cluster_with_minhash_signatures', fake_cluster)

    class Keep(BaseFilter):
        def __init__(self):
            self.seen = []

        def _filter(self, input):
            self.seen.append(len(input))
            return list(input)[:3]
    rng = np.random.default_rng(0)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    seqs = [letters[rng.integers(0, 4, 400)].tobytes().decode() for _ in range(5)] + ['ACGT' * 10]
    genomes = [[genome.Genome.from_one_seq(s) for s in seqs]]
    first, second = Keep(), Keep()
    pd = probe_designer.ProbeDesigner(genomes, [first, second], probe_length=100, probe_stride=50,
                                      seq_length_to_skip=50, cluster_threshold=0.15, cluster_merge_after=first,
                                      cluster_method='choose', cluster_fragment_length=150)
    pd.design()
    # 5 sequences of 400 nt in fragments of 150 (the last one flush with the end): 3 each; the 40-nt sequence is skipped
    assert calls == {'n': 15, 'method': 'hierarchical'}           # fragments shorter than the average sequence
    assert len(first.seen) == 2                                    # once per cluster
    assert len(second.seen) == 1 and 0 < second.seen[0] <= 6       # once, ungrouped, on the merged probes
    assert 0 < len(pd.final_probes) <= 3
    assert len(pd.candidate_probes) == sum(first.seen)


def test_filters_fail_loudly_without_a_gpu():
    """No CPU fallback: with no usable device the filter raises instead of computing elsewhere."""
    from catch_b200 import _lib, probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    from tests import helpers
    try:
        _lib.Context(0)
    except _lib.CatchB200Error:
        f = SetCoverFilter(0, 12, kmer_probe_map_k=5)
        with pytest.raises(_lib.CatchB200Error):
            f.filter([[probe.Probe.from_str('ACGTACGTACGT')]], helpers.to_genomes([[['ACGTACGTACGTACGT']]]),
                     input_is_grouped=True)
    with pytest.raises(NotImplementedError):
        SetCoverFilter(0, 10, custom_cover_range_fn=('x.py', 'f'))


def test_mt19937_replay_matches_numpy_sync_and_background():
    """cb_mt19937_randint (and its worker-thread variant) continue numpy's legacy stream exactly:
    same draws as np.random.randint and the same generator state afterwards (probe.py:393-396)."""
    from catch_b200 import _lib
    for bound, shape, seed in [(56, (1000, 20), 7), (81, (333, 20), 1), (1, (5, 20), 2), (64, (700, 20), 3),
                               (2, (10, 3), 4)]:
        np.random.seed(seed)
        np.random.randint(0, 10, size=17)             # start from a position inside a block
        state0 = np.random.get_state()
        want = np.random.randint(0, bound, size=shape)
        after = int(np.random.randint(0, 1 << 30))
        np.random.set_state(state0)
        got = _lib.legacy_randint(bound, shape)
        assert np.array_equal(got, want)
        assert int(np.random.randint(0, 1 << 30)) == after
        np.random.set_state(state0)
        pending = _lib.PendingRandint(bound, shape)
        got = pending.result()
        assert np.array_equal(got, want)
        assert int(np.random.randint(0, 1 << 30)) == after


def test_split_lengths():
    from catch_b200 import _lib
    strs = ['ACGT', '', 'A', 'GGGTTT', '']
    raw = '\n'.join(strs).encode()
    assert _lib.split_lengths(raw, len(strs)).tolist() == [len(s) for s in strs]
    assert _lib.split_lengths(raw, len(strs) + 1) is None
    assert _lib.split_lengths(raw, len(strs) - 1) is None
    assert _lib.split_lengths(b'', 0).tolist() == []
    assert _lib.split_lengths(b'', 1).tolist() == [0]


def test_gather_probes_c_helper_matches_python():
    """_fastpack.gather (one C pass over the Probe list) gives the bytes and lengths that the plain
    Python join would; wide characters are rejected like the Python path rejects them."""
    from catch_b200 import coverage as cov
    from catch_b200 import probe
    assert cov._fastpack is not None, "catch_b200/_fastpack.so is not built"
    strs = ['ACGT', '', 'N' * 300, 'ACGTNRYK', 'A']
    for items in (strs, [probe.Probe.from_str(s) for s in strs], tuple(strs), []):
        data, lens = cov.gather_probes(items)
        assert data == ''.join(strs[:len(items)]).encode() and lens.tolist() == [len(s) for s in strs[:len(items)]]
    saved, cov._fastpack = cov._fastpack, None
    try:
        data2, lens2 = cov.gather_probes([probe.Probe.from_str(s) for s in strs])
    finally:
        cov._fastpack = saved
    assert data2 == ''.join(strs).encode() and lens2.tolist() == [len(s) for s in strs]
    for bad in (['ACΔT'], [probe.Probe.from_str('A\U0001F600')]):
        with pytest.raises(ValueError):
            cov.gather_probes(bad)
    with pytest.raises((TypeError, AttributeError)):
        cov._fastpack.gather([1, 2], 'seq_str')
    # latin-1 characters are single bytes and pass through
    data, lens = cov.gather_probes(['A\xe9'])
    assert data == b'A\xe9' and lens.tolist() == [2]


def test_background_draw_can_be_cancelled_without_touching_the_rng():
    from catch_b200 import coverage as cov
    np.random.seed(11)
    before = np.random.get_state()[1].copy()
    drawn = cov.draw_seeds(np.full(5000, 75, dtype=np.int32), 2, 60, 20, background=True)
    cov.cancel_draw(drawn)
    assert np.array_equal(np.random.get_state()[1], before)
    k, seeds, mode = cov.finish_draw(cov.draw_seeds(np.full(5000, 75, dtype=np.int32), 2, 60, 20, background=True))
    np.random.seed(11)
    assert mode == 'random' and np.array_equal(seeds, np.random.randint(0, 56, size=(5000, 20)))


def test_gather_large_input_takes_the_threaded_copy():
    """Pass 2 of the gather runs on several threads with the GIL released above a size threshold (24 MB; the
    threshold is read once per process, so the threaded path is exercised in a child with CB_GATHER_PAR_MB=1)."""
    import subprocess
    code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from catch_b200 import coverage as cov
rng = np.random.default_rng(1)
strs = [''.join('ACGT'[i] for i in rng.integers(0, 4, int(n))) for n in rng.integers(50_000, 150_000, 40)]
strs += ['', 'N', 'ACGT' * 3]
assert sum(map(len, strs)) > (2 << 20)
data, lens = cov.gather_probes(strs)
assert data == ''.join(strs).encode() and lens.tolist() == [len(s) for s in strs]
print('ok')
''' % ROOT
    for mb in ('1', '100000'):
        r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, CB_GATHER_PAR_MB=mb), capture_output=True,
                           text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr[-1500:]


def test_mt19937_simd_and_scalar_paths_agree_with_numpy():
    """The AVX-512 loop (taken when the CPU has it) and the portable loop give numpy's draws and
    leave the generator at the same position, for sizes around the 16-lane and 624-word edges."""
    import ctypes as C
    from catch_b200 import _lib
    L = _lib.load()
    rs = np.random.RandomState(123)
    for trial in range(60):
        bound = int(rs.choice([2, 3, 56, 64, 81, 100, 255, 256, 257, 1000, 65536, 2 ** 31]))
        n = int(rs.choice([0, 1, 15, 16, 17, 31, 600, 623, 624, 625, 1250, 5000, 20001]))
        np.random.seed(trial)
        np.random.randint(0, 7, size=int(rs.randint(0, 700)))           # arbitrary position inside a block
        state0 = np.random.get_state()
        want = np.random.randint(0, bound, size=n)
        key_want, pos_want = np.random.get_state()[1].copy(), int(np.random.get_state()[2])
        for fn in (L.cb_mt19937_randint, L.cb_mt19937_randint_scalar):
            key = state0[1].astype(np.uint32).copy()
            pos = C.c_int32(int(state0[2]))
            out = np.full(n + 32, -1, dtype=np.int32)
            assert fn(key.ctypes.data, C.byref(pos), bound, n, out.ctypes.data) == 0
            assert np.array_equal(out[:n], want), (trial, bound, n)
            assert np.all(out[n + 16:] == -1)                             # nothing written far past the end
            # numpy refills lazily: compare the continuation of the stream, not the raw state
            np.random.set_state(('MT19937', key, pos.value, 0, 0.0))
            a = np.random.randint(0, 1 << 30, size=5)
            np.random.set_state(('MT19937', key_want, pos_want, 0, 0.0))
            assert np.array_equal(a, np.random.randint(0, 1 << 30, size=5)), (trial, bound, n)
        if bound <= 256:
            key = state0[1].astype(np.uint32).copy()
            pos = C.c_int32(int(state0[2]))
            out8 = np.full(n + 32, 7, dtype=np.uint8)
            assert L.cb_mt19937_randint_u8(key.ctypes.data, C.byref(pos), bound, n, out8.ctypes.data) == 0
            assert np.array_equal(out8[:n], want.astype(np.uint8))


def test_set_cover_drop_in_validates_like_the_reference():
    """The argument checks of approx_multiuniverse / approx (utils/set_cover.py:270-340, :88-95) come
    before any device work, so they can be exercised without a GPU."""
    from catch_b200.utils import set_cover as sc
    sets = {0: {0: {1, 2}}, 1: {0: {2, 3}}}
    with pytest.raises(ValueError, match="both arrays and IntervalSets"):
        sc.approx_multiuniverse(sets, use_arrays=True, use_intervalsets=True)
    with pytest.raises(ValueError, match="nonnegative"):
        sc.approx_multiuniverse(sets, costs={0: 1.0, 1: -2.0})
    with pytest.raises(ValueError, match="costs is missing"):
        sc.approx_multiuniverse(sets, costs={0: 1.0})
    with pytest.raises(ValueError, match="coverage fraction"):
        sc.approx_multiuniverse(sets, universe_p={0: 1.5})
    with pytest.raises(ValueError, match="universe_p is missing"):
        sc.approx_multiuniverse(sets, universe_p={7: 0.5})
    with pytest.raises(ValueError, match="ranks is missing"):
        sc.approx_multiuniverse(sets, ranks={0: 1})
    with pytest.raises(ValueError, match=r"p must be in \[0,1\]"):
        sc.approx({0: {1}}, p=-0.1)
    assert sc.approx_multiuniverse({}) == set()
    # elements -> intervals: runs of consecutive indices of the universe's element numbering
    index_of = {v: i for i, v in enumerate(['a', 'b', 'c', 'd', 'e'])}
    assert sc._as_intervals({'a', 'b', 'd'}, False, index_of) == [(0, 2), (3, 4)]
    assert sc._as_intervals((5, 9), True, None) == [(5, 9)]


def test_set_ordered_equals_set_of_probes():
    """near_duplicate_filter.set_ordered (set of sequences, C-level hashing) iterates in the order of the
    reference's set of Probe objects built by .add() in the same order (near_duplicate_filter.py:96-103)."""
    import random
    from catch_b200 import probe
    from catch_b200.filter.near_duplicate_filter import set_ordered
    rng = random.Random(4)
    for n in (0, 1, 5, 8, 9, 100, 5000, 70000):
        probes = [probe.Probe.from_str(''.join(rng.choice('ACGT') for _ in range(rng.choice([20, 40, 100]))))
                  for _ in range(n)]
        probes = list(dict.fromkeys(probes))
        want = set()
        for p in probes:
            want.add(p)
        got = set_ordered(probes)
        assert len(got) == len(probes) and all(a is b for a, b in zip(got, list(want)))
    dup = [probe.Probe.from_str('ACGT'), probe.Probe.from_str('ACGT'), probe.Probe.from_str('TTTT')]
    assert [p.seq_str for p in set_ordered(dup)] == [p.seq_str for p in list(set(dup))]


def test_probe_batch_tiling_equals_candidate_probes():
    """ProbeBatch.from_sequences (one strided buffer) yields exactly the probes of
    filter/candidate_probes.py:21-182 in the same order, N-run handling and flanking flags included."""
    import random
    from catch_b200.filter import candidate_probes as cp
    from catch_b200.probe_batch import ProbeBatch
    rng = random.Random(9)
    n_cases = 0
    for case in range(60):
        pl = rng.choice([20, 50, 100])
        ps = rng.choice([7, 10, 25, 50, 100, 130])
        seqs = []
        for _ in range(rng.randint(1, 4)):
            n = rng.randint(pl, 6 * pl + rng.randint(0, 40))
            s = [rng.choice('ACGT') for _ in range(n)]
            for _ in range(rng.choice([0, 0, 1, 3])):            # N runs of length 1..6, sometimes at the ends
                a = rng.choice([0, rng.randrange(n), n - 1])
                for i in range(a, min(n, a + rng.randint(1, 6))):
                    s[i] = 'N'
            seqs.append(''.join(s))
        skip = rng.choice([None, None, pl + 5])
        want = cp.make_candidate_probes_from_sequences(seqs, pl, ps, seq_length_to_skip=skip)
        got = ProbeBatch.from_sequences(seqs, pl, ps, seq_length_to_skip=skip)
        assert got is not None
        assert got.strs() == [p.seq_str for p in want]
        assert [p.is_flanking_n_string for p in got] == [p.is_flanking_n_string for p in want]
        assert [p.seq_str for p in got[1:4]] == [p.seq_str for p in want[1:4]]
        n_cases += len(want) > 0
    assert n_cases > 40
    assert ProbeBatch.from_sequences(['ACGT'], 10, 5) is None        # shorter than a probe: per-object path
    b = ProbeBatch.from_sequences(['ACGTACGTAC', 'TTTTTTTTTTTT'], 10, 5)
    assert b.take([1, 0]).strs() == ['TTTTTTTTTT', 'ACGTACGTAC'] and len(b) == 3


def test_cluster_host_logic_on_the_reference_test_vectors():
    """The toy distance functions of the reference's catch/utils/tests/test_cluster.py (its expected values restated
    as data): connected components, condensed matrix layout, hierarchical clustering.  No device involved: a plain
    callable is the caller's own distance function."""
    from catch_b200.utils import cluster
    seqs = ['a', 'b', 'x', 'm', 'c', 'o', 'n', 'z', 'w', 'y', 'v', 'd', 'k']

    def dist(i, j):
        return abs(ord(seqs[i]) - ord(seqs[j]))
    assert cluster.find_connected_components(len(seqs), dist, 1) == [[2, 7, 8, 9, 10], [0, 1, 4, 11], [3, 5, 6], [12]]
    seqs = ['a', 'c', 'b', 'c']
    assert cluster.find_connected_components(4, dist, 1) == [[0, 1, 2, 3]]
    seqs = ['a', 'z', 'm']
    assert sorted(cluster.find_connected_components(3, dist, 1)) == [[0], [1], [2]]
    assert cluster.find_connected_components(0, dist, 1) == []
    m2 = np.array([[0, 1, 100], [1, 0, 2], [100, 2, 0]])
    cond = cluster.create_condensed_dist_matrix(3, lambda i, j: m2[i][j])
    assert cond.dtype == np.float32 and cond.tolist() == [1.0, 100.0, 2.0]
    assert cluster.cluster_hierarchically_from_dist_matrix(cond, 10) == [[0, 1], [2]]
    d = {(0, 1): 100, (0, 2): 100, (1, 2): 1}
    cond = cluster.create_condensed_dist_matrix(3, lambda i, j: d[(i, j)])
    assert cluster.cluster_hierarchically_from_dist_matrix(cond, 10) == [[1, 2], [0]]
    assert cluster.cluster_hierarchically_from_dist_matrix(cluster.create_condensed_dist_matrix(1, None), 10) == [[0]]
    assert cluster._jaccard_dist_from_mash_dist(0.1, 12) == 1.0 - 1.0 / (2.0 * np.exp(12 * 0.1) - 1)


def test_fasta_reader_matches_reference_fixtures(tmp_path):
    """the native FASTA parser (and the Python line loop behind it) against outputs of the reference's
    seq_io.read_fasta recorded by tests/golden/make_f4_golden.py."""
    import base64
    import gzip
    from collections import OrderedDict
    from catch_b200.utils import seq_io
    from tests import golden_io
    assert seq_io._fastpack is not None and hasattr(seq_io._fastpack, 'parse_fasta')
    n_ok = 0
    for c in golden_io.load('f4_reference.json.gz'):
        fn = str(tmp_path / ('t.fasta.gz' if c['gz'] else 't.fasta'))
        with (gzip.open(fn, 'wb') if c['gz'] else open(fn, 'wb')) as f:
            f.write(base64.b64decode(c['data']))
        for reader in (seq_io.read_fasta, seq_io._read_fasta_lines):
            if c['out'] == 'AssertionError':
                with pytest.raises(AssertionError):
                    reader(fn, **c['kw'])
            else:
                got = reader(fn, **c['kw'])
                assert isinstance(got, OrderedDict)
                assert [[n, s] for n, s in got.items()] == c['out']
                n_ok += 1
    assert n_ok > 300
    # non-ASCII bytes take the line loop; one Genome per record
    fn = str(tmp_path / 'u.fasta')
    with open(fn, 'w', encoding='utf-8') as f:
        f.write('>g\u00e9nome\nacgt-y\n>b\nAC\nGT\n')
    assert list(seq_io.read_fasta(fn).items()) == [('g\u00e9nome', 'ACGTN'), ('b', 'ACGT')]
    assert [g.seqs for g in seq_io.read_genomes_from_fasta(fn)] == [['ACGTN'], ['ACGT']]


def test_probe_list_fingerprint_sees_order_and_content():
    """The checksum ranks compare before sharding ONE probe list by position (catch_b200/coverage.py: fingerprint)."""
    import ctypes
    from catch_b200 import coverage as cov, probe
    ps = [probe.Probe('ACGT' * 20 + str(i % 10) * 5) for i in range(1000)]
    g = cov.gather_probes(ps)
    assert cov.fingerprint(g) == cov.fingerprint(cov.gather_probes(list(ps)))
    assert cov.fingerprint(g) != cov.fingerprint(cov.gather_probes(ps[::-1]))
    assert cov.fingerprint(g) != cov.fingerprint(cov.gather_probes(ps[:-1]))
    buf = ctypes.create_string_buffer(g[0], len(g[0]))                      # the same bytes behind an address
    assert cov.fingerprint((ctypes.addressof(buf), g[1])) == cov.fingerprint(g)
    assert cov.fingerprint((b'', np.zeros(0, np.int32))) == 0


def test_set_cover_filter_survives_pickling():
    """The filter object keeps per-thread state and a device context; neither travels through pickle / copy."""
    import copy
    import pickle
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    f = SetCoverFilter(mismatches=2, lcf_thres=60, cover_extension=10, coverage=0.9)
    for g in (pickle.loads(pickle.dumps(f)), copy.deepcopy(f)):
        assert (g.mismatches, g.lcf_thres, g.cover_extension, g.coverage) == (2, 60, 10, 0.9)
        assert g._ctx is None and getattr(g._tls, 'ctx', None) is None


def test_draw_replay_reproduces_the_sequential_stream():
    """_DrawReplay (every rank / worker gets the seed draws of its groupings from one helper thread that replays ALL
    groupings in order) against plain sequential draws: same draws for every grouping, same RNG state afterwards;
    lists of Probe objects, str lists, ProbeBatch input, a list whose length samples disagree (measured in full) and
    one whose odd-length probes the samples miss (the owner's check raises)."""
    from catch_b200 import coverage as cov, probe
    from catch_b200.filter import set_cover_filter as scf_mod
    from catch_b200.probe_batch import ProbeBatch
    rng = np.random.default_rng(3)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)

    def strs(n, L):
        flat = letters[rng.integers(0, 4, (n, L), dtype=np.uint8)].tobytes().decode()
        return [flat[i * L:(i + 1) * L] for i in range(n)]
    g0 = [probe.Probe.from_str(s) for s in strs(300, 100)]
    g1 = strs(120, 100)
    g2 = ProbeBatch(letters[rng.integers(0, 4, (200, 100), dtype=np.uint8)])
    g3 = [probe.Probe.from_str(s) for s in (strs(50, 100) + strs(50, 80) + strs(60, 100))]      # samples see both lengths
    g4 = []
    inputs = [g0, g1, g2, g3, g4]
    flt = scf_mod.SetCoverFilter(mismatches=2, lcf_thres=60)

    def lengths(g):
        return cov.probe_lengths(g) if len(g) else np.zeros(0, dtype=np.int32)
    np.random.seed(11)
    want = []
    for g in inputs:
        lens = lengths(g)
        want.append(np.asarray(cov.finish_draw(cov.draw_seeds(lens, 2, 60, 20))[1]) if len(lens) else None)
    after = int(np.random.randint(0, 1 << 30))
    for owner, rank in (([0] * 5, 0), ([0, 1, 0, 1, 0], 1), ([1, 1, 1, 1, 1], 0)):
        np.random.seed(11)
        rep = scf_mod._DrawReplay(flt, inputs, owner, rank)
        for g in range(5):
            if owner[g] != rank:
                continue
            drawn, drawn_tol = rep.get(g, lengths(inputs[g]) if len(inputs[g]) else None)
            assert drawn_tol is None
            if want[g] is None:
                assert drawn is None
            else:
                assert np.array_equal(np.asarray(cov.finish_draw(drawn)[1]), want[g]), g
        rep.finish()
        assert int(np.random.randint(0, 1 << 30)) == after
    # odd lengths that the 48 samples miss: the owner's check raises
    odd = strs(400, 100)
    odd[7] = odd[7][:80]
    np.random.seed(11)
    rep = scf_mod._DrawReplay(flt, [odd], [0], 0)
    with pytest.raises(scf_mod._LengthsNotUniform):
        rep.get(0, cov.probe_lengths(odd))
    rep.finish()


def test_connected_components_search_equals_the_reference_walk():
    """find_connected_components with a device-style distance provider (rows thresholded to neighbour lists, fetched
    ahead in batches) against the oracle's plain restatement of the reference's walk over every pair: random planar
    point sets, ties at both thresholds, the early-stop heuristic on and off.  The provider here is a numpy stand-in
    for cb_sketch_near_rows, so the host logic is checked without a GPU."""
    from catch_b200.utils import cluster
    from oracle import oracle

    class FakeSketches(cluster.SketchSet):
        def __init__(self, D):
            self.D, self.n, self._row_cache, self.calls = D, len(D), {}, 0
            self.h = None

        @property
        def ctx(self):
            return self

        def sketch_near_rows(self, h, want, thr):
            self.calls += 1
            off, idx, dist = [0], [], []
            for j in want:
                c = np.flatnonzero(self.D[j] <= thr)
                idx.append(c.astype(np.uint32))
                dist.append(self.D[j][c])
                off.append(off[-1] + len(c))
            return np.array(off, dtype=np.int64), np.concatenate(idx), np.concatenate(dist)

    rng = np.random.default_rng(0)
    visited = calls = 0
    for trial in range(150):
        n = int(rng.integers(1, 250))
        pts = rng.random((n, 2)) * rng.choice([1, 3, 10])
        D = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1))
        if trial % 3 == 0:
            D = np.round(D, 1)                       # many distances exactly at a threshold
        thr = float(rng.choice([0.1, 0.3, 0.6]))
        early = float(rng.choice([0, 0.05, 0.1, 0.3]))
        f = FakeSketches(D)
        got = cluster.find_connected_components(n, f, thr, early)
        want = oracle.connected_components(n, lambda i, j: D[i, j], thr, early)
        assert got == want, trial
        visited += n
        calls += f.calls
    assert calls < visited / 3                        # rows are fetched in batches, not one call per vertex


def test_fasta_reader_multi_span_path_matches_reference_fixtures(tmp_path):
    """The native parser cuts large files at header lines and parses the spans on several threads; with the size
    threshold at one byte (read once per process: a child interpreter) the 400 reference-recorded files go through
    that path, cuts included."""
    import subprocess
    code = r'''
import base64, gzip, os, sys
sys.path.insert(0, %r)
from catch_b200.utils import seq_io
from tests import golden_io
tmp = %r
n_ok = n_rejected = 0
for c in golden_io.load('f4_reference.json.gz'):
    fn = os.path.join(tmp, 't.fasta.gz' if c['gz'] else 't.fasta')
    with (gzip.open(fn, 'wb') if c['gz'] else open(fn, 'wb')) as f:
        f.write(base64.b64decode(c['data']))
    try:
        got = [[n, s] for n, s in seq_io.read_fasta(fn, **c['kw']).items()]
    except AssertionError:
        got = 'AssertionError'
    assert got == c['out'], c
    n_ok += got != 'AssertionError'
    n_rejected += got == 'AssertionError'
print(n_ok, n_rejected)
''' % (ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, CB_FASTA_PAR_BYTES='1'), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    n_ok, n_rejected = map(int, r.stdout.split())
    assert n_ok > 150 and n_rejected > 50


def test_fasta_streaming_reader_matches_reference_fixtures(tmp_path):
    """seq_io.iterate_fasta (blocks cut at line ends, scanned natively) against the reference's iterate_fasta outputs
    recorded by tests/golden/make_f4_golden.py, with block sizes from a few bytes (every cut position) to one block."""
    import base64
    import gzip
    from catch_b200.utils import seq_io
    from tests import golden_io
    assert hasattr(seq_io._fastpack, 'fasta_stream_block')
    n = 0
    for k, c in enumerate(golden_io.load('f4_reference.json.gz')):
        fn = str(tmp_path / ('t.fasta.gz' if c['gz'] else 't.fasta'))
        with (gzip.open(fn, 'wb') if c['gz'] else open(fn, 'wb')) as f:
            f.write(base64.b64decode(c['data']))
        for bs in (1 + k % 9, 50 + k % 40, 64 << 20):
            got = list(seq_io.iterate_fasta(fn, replace_degenerate=c['kw']['replace_degenerate'], block_bytes=bs))
            assert got == c['iter'], (k, bs)
            n += len(got)
    assert n > 1000
    fn = str(tmp_path / 'u.fasta')
    with open(fn, 'w', encoding='utf-8') as f:                                # non-ASCII: the line loop takes over
        f.write('>a\nACGY\n>gé\nACéT\nyy\n')
    for bs in (3, 64 << 20):
        assert list(seq_io.iterate_fasta(fn, block_bytes=bs)) == ['ACGN', 'ACéTyy']
