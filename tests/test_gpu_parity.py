"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200."""
import os
import random

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import oracle
    return oracle


def _device_quads(ctx, probe_strs, genomes, params, k=None):
    from catch_b200 import coverage as cov
    group = cov.PackedGroup(ctx, probe_strs, genomes)
    plan = cov.SeedPlan(probe_strs, params['mismatches'], params['lcf_thres'], params['kmer_probe_map_k'])
    cover, st = cov.compute_cover(ctx, group, plan, params['mismatches'], params['lcf_thres'],
                                  params['island_of_exact_match'], params['cover_extension'])
    pid, gen, s, e = ctx.cover_export(cover)
    group.free()
    quads = np.stack([pid, gen.astype(np.int64), s, e], axis=1) if len(pid) else np.zeros((0, 4), np.int64)
    return quads, cover, st


@pytest.mark.parametrize('case', range(24))
def test_coverage_and_setcover_match_oracle(ctx, case):
    O = _oracle()
    alphabet = 'ACGT' if case % 4 else 'ACGTRYKMSW'
    groups, cands, params = helpers.random_case(case, alphabet=alphabet)
    for probe_strs, genomes in zip(cands, groups):
        if not probe_strs:
            continue
        seed = 1000 + case
        np.random.seed(seed)
        k, seeds, _ = O.choose_seeds(probe_strs, params['mismatches'], params['lcf_thres'],
                                     min_k=params['kmer_probe_map_k'], k=params['kmer_probe_map_k'])
        sm = O.SeedMap(probe_strs, seeds, k)
        want = O.make_sets_quads(sm, genomes, params['mismatches'], params['lcf_thres'],
                                 params['island_of_exact_match'], params['cover_extension'])
        np.random.seed(seed)
        got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.array_equal(got, want)
        # stage B on the same cover, both modes of the greedy kernel
        cov_p = params['coverage']
        if cov_p <= 1.0:
            up = np.full(len(genomes), float(cov_p))
        else:
            up = np.array([float(min(cov_p, sum(map(len, g)))) / sum(map(len, g)) for g in genomes])
        want_picks = O.set_cover_quads(want, len(probe_strs), len(genomes), None, up, None)
        picks, st_b = ctx.setcover(cover, len(probe_strs), None, up)
        assert picks.tolist() == want_picks
        cover.free()


@pytest.mark.parametrize('case', range(12))
def test_filter_end_to_end_matches_oracle(ctx, case):
    """SetCoverFilter.filter() vs the oracle's restatement of the reference filter, including
    the order of the returned probes."""
    O = _oracle()
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    groups, cands, params = helpers.random_case(100 + case)
    genomes = helpers.to_genomes(groups)
    probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
    f = SetCoverFilter(**params)
    np.random.seed(7 + case)
    random.seed(7 + case)
    out = f.filter(probes, genomes, input_is_grouped=True)
    np.random.seed(7 + case)
    random.seed(7 + case)
    want = O.set_cover_filter(cands, groups, params['mismatches'], params['lcf_thres'],
                              params['island_of_exact_match'], params['coverage'],
                              params['cover_extension'], params['kmer_probe_map_k'])
    for g, (o, w) in enumerate(zip(out, want)):
        ids = {id(p): i for i, p in enumerate(probes[g])}
        assert [ids[id(p)] for p in o] == w


def test_ranks_and_full_mode(ctx):
    """Ranks (set_cover.py:349,491,522-526) and the p<1 clamp on a shared cover."""
    O = _oracle()
    rng = random.Random(5)
    groups, cands, params = helpers.random_case(3)
    probe_strs, genomes = cands[0], groups[0]
    np.random.seed(11)
    got, cover, _ = _device_quads(ctx, probe_strs, genomes, params)
    for trial in range(6):
        ranks = np.array([rng.choice([0, 0, 1, 5]) for _ in probe_strs], dtype=np.int32)
        up = np.array([rng.choice([1.0, 0.9, 0.5, 0.1]) if trial % 2 else 1.0 for _ in genomes])
        want = O.set_cover_quads(got, len(probe_strs), len(genomes), None, up, ranks)
        picks, _ = ctx.setcover(cover, len(probe_strs), ranks, up)
        assert picks.tolist() == want
    cover.free()


def test_empty_inputs(ctx):
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    f = SetCoverFilter(0, 20)
    assert f.filter([[]], [[]], input_is_grouped=True) == [[]]
    # a sequence shorter than k yields no coverage (probe.py:1204-1212)
    g = helpers.to_genomes([[['ACGTACGTAC']]])
    p = [[probe.Probe.from_str('ACGTACGTACGTACGTACGTACGTA')]]
    assert f.filter(p, g, input_is_grouped=True) == [[]]


@pytest.mark.parametrize('pl,m,lcf', [(75, 0, 75), (100, 0, 100), (75, 2, 75), (100, 2, 100), (100, 1, 100),
                                      (75, 3, 60), (100, 5, 30), (128, 2, 50), (120, 4, 120)])
def test_fast_path_probe_lengths(ctx, pl, m, lcf):
    """Probes of 65..128 nt take the specialised 128-bit scan path for k in {20,25,50,75,100} and
    the generic one otherwise; both must reproduce the oracle's intervals exactly (N, multi-sequence
    genomes, probes hanging off both ends, cover extension)."""
    O = _oracle()
    rng = random.Random(pl * 100 + m)
    groups = helpers.random_groups(rng, n_groups=1, anc_len=(400, 900), max_genomes=8)
    genomes = groups[0]
    seqs = [s for g in genomes for s in g]
    probe_strs = [x for x in dict.fromkeys(helpers.tile_candidates(seqs, pl, 20)) if len(x) >= 20]
    # probes that overhang: built from sequence ends plus random tails
    for s in seqs[:4]:
        if len(s) >= 40:
            tail = ''.join(rng.choice('ACGT') for _ in range(pl - 40))
            probe_strs.append((tail + s[:40])[:pl])
            probe_strs.append((s[-40:] + tail)[:pl])
    params = dict(mismatches=m, lcf_thres=lcf, island_of_exact_match=rng.choice([0, 0, 15]),
                  cover_extension=rng.choice([0, 50]), kmer_probe_map_k=20)
    np.random.seed(pl + m)
    k, seeds, mode = O.choose_seeds(probe_strs, m, lcf, min_k=20, k=20)
    sm = O.SeedMap(probe_strs, seeds, k)
    want = O.make_sets_quads(sm, genomes, m, lcf, params['island_of_exact_match'], params['cover_extension'])
    np.random.seed(pl + m)
    got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
    cover.free()
    assert len(want) > 0
    assert np.array_equal(got, want), (k, mode)


@pytest.mark.parametrize('pl,m,lcf,alphabet', [
    (150, 3, 100, 'ACGT'), (200, 0, 200, 'ACGT'), (256, 8, 120, 'ACGT'), (190, 10, 60, 'ACGTN'),
    (64, 1, 40, 'ACGT'), (129, 2, 129, 'ACGT'),
    (90, 2, 50, 'ABCDEFGHIJKLMNOPQRSTUVWXYZ'), (40, 1, 30, 'ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789')])
def test_generic_path_long_probes_and_wide_alphabets(ctx, pl, m, lcf, alphabet):
    """Probes up to CB_MAX_PROBE_LEN (3-4 mask words) and alphabets needing 5-6 bit planes take the
    generic multi-word path; intervals must still match the oracle exactly."""
    O = _oracle()
    rng = random.Random(pl + 7 * m)
    groups = helpers.random_groups(rng, alphabet=alphabet, n_groups=1, anc_len=(600, 1200), max_genomes=6,
                                   with_n=('N' in alphabet or len(alphabet) > 4))
    genomes = groups[0]
    seqs = [s for g in genomes for s in g]
    probe_strs = [x for x in dict.fromkeys(helpers.tile_candidates(seqs, pl, 37)) if len(x) >= 20]
    params = dict(mismatches=m, lcf_thres=lcf, island_of_exact_match=rng.choice([0, 12]),
                  cover_extension=rng.choice([0, 30]), kmer_probe_map_k=20)
    np.random.seed(pl)
    k, seeds, mode = O.choose_seeds(probe_strs, m, lcf, min_k=20, k=20)
    sm = O.SeedMap(probe_strs, seeds, k)
    want = O.make_sets_quads(sm, genomes, m, lcf, params['island_of_exact_match'], params['cover_extension'])
    np.random.seed(pl)
    got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
    assert len(want) > 0
    assert np.array_equal(got, want), (k, mode)
    up = np.full(len(genomes), 1.0)
    assert ctx.setcover(cover, len(probe_strs), None, up)[0].tolist() == \
        O.set_cover_quads(want, len(probe_strs), len(genomes), None, up, None)
    cover.free()


def test_limits_are_reported(ctx):
    """Outside the supported envelope the library says so instead of computing something else."""
    from catch_b200 import _lib, probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    g = helpers.to_genomes([[['ACGT' * 100]]])
    with pytest.raises(_lib.CatchB200Error):          # probe longer than 256 nt
        SetCoverFilter(0, 300, kmer_probe_map_k=20).filter([[probe.Probe.from_str('ACGT' * 75)]], g,
                                                           input_is_grouped=True)
    with pytest.raises(_lib.CatchB200Error):          # more than 31 mismatches
        SetCoverFilter(40, 30).filter([[probe.Probe.from_str('ACGT' * 25)]], g, input_is_grouped=True)


@pytest.mark.parametrize('mode,cap', [('par', None), ('par', '1'), ('par', '3'), ('legacy', None)])
def test_greedy_kernels_agree_with_oracle(ctx, monkeypatch, mode, cap):
    """Both greedy kernels (parallel rounds, csrc/rounds.cu; one pick per iteration, csrc/setcover.cu) give
    the oracle's pick SEQUENCE; a tiny candidate list forces the overflow and rebuild paths of the
    parallel-rounds kernel."""
    O = _oracle()
    monkeypatch.setenv('CB_GREEDY', mode)
    if cap:
        monkeypatch.setenv('CB_GREEDY_LIST_CAP', cap)
    rng = random.Random(17)
    for case in (1, 2, 5, 7, 9):
        groups, cands, params = helpers.random_case(case)
        probe_strs, genomes = cands[0], groups[0]
        if not probe_strs:
            continue
        np.random.seed(3)
        got, cover, _ = _device_quads(ctx, probe_strs, genomes, params)
        for ranks in (None, np.array([rng.choice([0, 0, 2, 7]) for _ in probe_strs], dtype=np.int32)):
            want = O.set_cover_quads(got, len(probe_strs), len(genomes), None, None, ranks)
            picks, st = ctx.setcover(cover, len(probe_strs), ranks, None)
            assert picks.tolist() == want
        cover.free()


def test_parallel_rounds_at_scale(ctx, monkeypatch):
    """Size-independent check at a size the oracle would take minutes for: the parallel-rounds
    kernel and the one-pick-per-iteration kernel give the same pick sequence, and the selection
    covers every universe bit."""
    from catch_b200 import coverage as cov
    seqs = helpers.synthetic_genomes(120, 6000, 0.03, seed=5)
    cands = list(dict.fromkeys(helpers.tile_candidates(seqs, 75, 25)))
    group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
    np.random.seed(7)
    plan = cov.SeedPlan(cands, 2, 60, 20)
    cover, _ = cov.compute_cover(ctx, group, plan, 2, 60, 0, 50)
    group.free()
    res = {}
    for mode in ('par', 'legacy'):
        monkeypatch.setenv('CB_GREEDY', mode)
        res[mode], st = ctx.setcover(cover, len(cands), None, None)
    assert len(res['par']) > 50
    assert res['par'].tolist() == res['legacy'].tolist()
    pid, gen, s, e = ctx.cover_export(cover)
    cover.free()
    chosen = np.zeros(len(cands), dtype=bool)
    chosen[res['par']] = True
    L = 6000
    allc = np.zeros(len(seqs) * L + 1, dtype=np.int32)
    selc = np.zeros(len(seqs) * L + 1, dtype=np.int32)
    np.add.at(allc, gen * L + s, 1)
    np.add.at(allc, gen * L + e, -1)
    m = chosen[pid]
    np.add.at(selc, gen[m] * L + s[m], 1)
    np.add.at(selc, gen[m] * L + e[m], -1)
    assert np.array_equal(np.cumsum(allc) > 0, np.cumsum(selc) > 0)


def test_merge_paths_small_and_large_range_lists(ctx):
    """Probes with a handful of ranges take the warp-per-probe merge, probes with more than 1024
    ranges (here: ~1300 near-identical genomes) the block-per-probe merge; both against the oracle."""
    O = _oracle()
    rng = random.Random(23)
    anc = ''.join(rng.choice('ACGT') for _ in range(260))
    anc_b = ''.join(rng.choice('ACGT') for _ in range(260))
    genomes = [[helpers.mutate(rng, anc, 0.01)] for _ in range(1300)] + \
              [[helpers.mutate(rng, anc_b, 0.01)] for _ in range(4)]
    probe_strs = list(dict.fromkeys(helpers.tile_candidates([g[0] for g in genomes[:6] + genomes[-2:]], 60, 20)))
    params = dict(mismatches=3, lcf_thres=40, island_of_exact_match=0, cover_extension=5, kmer_probe_map_k=15)
    np.random.seed(5)
    k, seeds, _ = O.choose_seeds(probe_strs, 3, 40, min_k=15, k=15)
    want = O.make_sets_quads(O.SeedMap(probe_strs, seeds, k), genomes, 3, 40, 0, 5)
    np.random.seed(5)
    got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
    counts = np.bincount(want[:, 0], minlength=len(probe_strs))
    assert counts.max() > 1024 and counts.min() < 512
    assert np.array_equal(got, want)
    picks, _ = ctx.setcover(cover, len(probe_strs), None, None)
    assert picks.tolist() == O.set_cover_quads(want, len(probe_strs), len(genomes))
    cover.free()


def test_parallel_rounds_tie_heavy_regime(ctx, monkeypatch):
    """Exact-match parameters (m=0, lcf = probe length) make thousands of probes tie at the same small
    gain and nearly every probe gets picked: the regime that exercises the histogram-chosen list
    threshold, the id-level tie handling and many winners per round.  The parallel-rounds kernel must
    give the one-pick kernel's sequence, every time."""
    from catch_b200 import coverage as cov
    gens = helpers.synthetic_influenza(40, seed=3)
    seqs = [seg for g in gens for seg in g]
    cands = list(dict.fromkeys(helpers.tile_candidates(seqs, 100, 50)))
    group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
    np.random.seed(7)
    plan = cov.SeedPlan(cands, 0, 100, 20)
    assert plan.mode == 'pigeonhole'
    cover, _ = cov.compute_cover(ctx, group, plan, 0, 100, 0, 50)
    group.free()
    monkeypatch.setenv('CB_GREEDY', 'legacy')
    want, _ = ctx.setcover(cover, len(cands), None, None)
    assert len(want) > 1000
    for cap in ('64', '2048', '2048'):
        monkeypatch.setenv('CB_GREEDY', 'par')
        monkeypatch.setenv('CB_GREEDY_LIST_CAP', cap)
        got, st = ctx.setcover(cover, len(cands), None, None)
        assert got.tolist() == want.tolist()
    cover.free()


def test_duplicate_filter_matches_ordered_dict(ctx):
    """DuplicateFilter on the device == list(OrderedDict.fromkeys(input)) (duplicate_filter.py:20-26):
    first occurrences, input order, the same Probe objects; mixed lengths, N, one-symbol alphabets."""
    from collections import OrderedDict
    from catch_b200 import probe
    from catch_b200.filter.duplicate_filter import DuplicateFilter
    rng = random.Random(3)
    f = DuplicateFilter()
    f._ctx = ctx
    for alphabet, n, lens in (('ACGT', 5000, (20, 21, 40)), ('ACGTN', 3000, (75,)), ('A', 50, (1, 2, 3)),
                              ('ACGTRYKM', 2000, (100, 130, 256))):
        pool = [''.join(rng.choice(alphabet) for _ in range(rng.choice(lens))) for _ in range(max(2, n // 4))]
        probes = [probe.Probe.from_str(rng.choice(pool)) for _ in range(n)]
        got = f.filter(probes)
        want = list(OrderedDict.fromkeys(probes))
        assert len(got) == len(want) and all(a is b for a, b in zip(got, want))
    assert f.filter([]) == []
    # grouped input goes through BaseFilter's per-grouping loop
    probes = [probe.Probe.from_str(s) for s in ('ACGT', 'ACGT', 'AC', 'ACGT', 'AC', 'T')]
    out = f.filter([probes, probes[2:]], input_is_grouped=True)
    assert [[p.seq_str for p in g] for g in out] == [['ACGT', 'AC', 'T'], ['AC', 'ACGT', 'T']]


def test_baseline_config2_full_size_properties(ctx, monkeypatch):
    """BASELINE config 2 at its full size (500 x 11 kb, -pl 75 -m 2 -l 60 -e 50; the oracle would need
    minutes): size-independent properties of the result.  (1) the parallel-rounds kernel gives the
    one-pick kernel's pick sequence; (2) the selection covers every universe bit that any candidate
    covers; (3) the picks are distinct and gains at pick time never increase along the sequence;
    (4) SetCoverFilter.filter() on host objects returns exactly those probes."""
    from catch_b200 import coverage as cov
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    seqs = helpers.synthetic_genomes(500, 11000, 0.03, seed=2)
    cands = list(dict.fromkeys(helpers.tile_candidates(seqs, 75, 50)))
    group = cov.PackedGroup(ctx, cands, [[s] for s in seqs])
    np.random.seed(7)
    plan = cov.SeedPlan(cands, 2, 60, 20)
    cover, st = cov.compute_cover(ctx, group, plan, 2, 60, 0, 50)
    group.free()
    picks = {}
    for mode in ('par', 'legacy'):
        monkeypatch.setenv('CB_GREEDY', mode)
        picks[mode], _ = ctx.setcover(cover, len(cands), None, None)
    monkeypatch.delenv('CB_GREEDY')
    assert picks['par'].tolist() == picks['legacy'].tolist()
    sel = picks['par']
    assert len(set(sel.tolist())) == len(sel) > 100
    pid, gen, s, e = ctx.cover_export(cover)
    cover.free()
    L = 11000
    lo, hi = gen * L + s, gen * L + e

    def covered(mask):
        n = len(seqs) * L + 1
        d = np.bincount(lo[mask], minlength=n) - np.bincount(hi[mask], minlength=n)
        return np.cumsum(d)[:-1] > 0
    chosen = np.zeros(len(cands), dtype=bool)
    chosen[sel] = True
    universe = covered(np.ones(len(pid), dtype=bool))
    assert np.array_equal(universe, covered(chosen[pid]))
    # gain of every pick at the time it is picked: new bits it adds, replayed on the host
    done = np.zeros(len(seqs) * L, dtype=bool)
    order = np.argsort(pid, kind='stable')
    starts = np.searchsorted(pid[order], np.arange(len(cands) + 1))
    prev = None
    for p in sel.tolist():
        gain = 0
        for i in order[starts[p]:starts[p + 1]]:
            seg = done[lo[i]:hi[i]]
            gain += int(seg.size - seg.sum())
            seg[:] = True
        assert gain > 0 and (prev is None or gain <= prev)
        prev = gain
    f = SetCoverFilter(mismatches=2, lcf_thres=60, cover_extension=50)
    f._ctx = ctx
    np.random.seed(7)
    out = f.filter([[probe.Probe.from_str(c) for c in cands]], helpers.to_genomes([[[x] for x in seqs]]),
                   input_is_grouped=True)
    assert sorted(p.seq_str for p in out[0]) == sorted(cands[i] for i in sel.tolist())


def test_setcover_with_costs_matches_oracle(ctx):
    """approx_multiuniverse with per-set float costs (utils/set_cover.py:426: min cost/gain, smallest id
    on ties), with and without ranks and partial cover: pick SEQUENCE against the oracle."""
    O = _oracle()
    rng = random.Random(41)
    for case in (1, 2, 5, 7):
        groups, cands, params = helpers.random_case(case)
        probe_strs, genomes = cands[0], groups[0]
        if not probe_strs:
            continue
        np.random.seed(3)
        got, cover, _ = _device_quads(ctx, probe_strs, genomes, params)
        for trial in range(4):
            costs = np.array([rng.choice([1.0, 1.0, 0.5, 2.0, 3.25, 0.1]) for _ in probe_strs])
            ranks = None if trial % 2 == 0 else np.array([rng.choice([0, 0, 3]) for _ in probe_strs], dtype=np.int32)
            up = None if trial < 2 else np.array([rng.choice([1.0, 0.7, 0.3]) for _ in genomes])
            want = O.set_cover_quads(got, len(probe_strs), len(genomes), costs, up, ranks)
            picks, _ = ctx.setcover(cover, len(probe_strs), ranks, up, costs=costs)
            assert picks.tolist() == want
        cover.free()


def test_probe_batch_filter_chain_equals_probe_lists(ctx):
    """candidates kept as one buffer (catch_b200/probe_batch.py) through DuplicateFilter /
    NearDuplicateFilter / SetCoverFilter give exactly the output of the same chain on lists of Probe objects
    (same probes, same order); Probe objects only appear for the selected ones."""
    from catch_b200 import probe
    from catch_b200.filter.duplicate_filter import DuplicateFilter
    from catch_b200.filter.near_duplicate_filter import NearDuplicateFilterWithMinHash
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    from catch_b200.probe_batch import ProbeBatch
    groups = [helpers.synthetic_genomes(25, 2500, 0.04, seed=21), helpers.synthetic_genomes(12, 1800, 0.02, seed=22)]
    groups[0][3] = groups[0][3][:700] + 'NNNN' + groups[0][3][704:]            # an N run: flanking probes
    genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
    batches = [ProbeBatch.from_sequences(g, 100, 50) for g in groups]
    lists = [[probe.Probe.from_str(s) for s in b.strs()] for b in batches]
    for first in ('dup', 'minhash'):
        outs = []
        for inp in (lists, batches):
            f1 = DuplicateFilter() if first == 'dup' else NearDuplicateFilterWithMinHash(0.5)
            f2 = SetCoverFilter(mismatches=3, lcf_thres=60, cover_extension=20)
            f1._ctx = f2._ctx = ctx
            np.random.seed(5)
            random.seed(5)
            mid = f1.filter(inp, genomes, input_is_grouped=True)
            if inp is batches:
                assert all(isinstance(m, ProbeBatch) for m in mid)
            out = f2.filter(mid, genomes, input_is_grouped=True)
            outs.append(([[p.seq_str for p in g] for g in mid], [[p.seq_str for p in g] for g in out]))
        if first == 'dup' or os.environ.get('PYTHONHASHSEED') == '0':
            assert outs[0] == outs[1]
        else:
            # the near-duplicate filter's output ORDER follows the interpreter's string hash seed on both
            # paths alike (list(set(...))); both are in THIS process, so the orders agree here too
            assert outs[0] == outs[1]
        assert sum(map(len, outs[0][1])) > 10


def test_pipelining_and_prefetch_do_not_change_the_output(ctx, monkeypatch):
    """Several groupings in one filter() call: two groupings at a time on two contexts with the draws replayed by a
    helper thread (default), one at a time with the helper thread that gathers the next grouping, and the plain loop
    all give the same selection in the same order, and leave numpy's RNG in the same state."""
    import random as pyrandom
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    from catch_b200.probe_batch import ProbeBatch
    groups = helpers.synthetic_taxa(7, 12, seed=9, length_range=(1500, 3000))
    genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
    cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
    probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
    batches = [ProbeBatch(np.frombuffer(''.join(c).encode(), dtype=np.uint8).reshape(len(c), 100)) for c in cands]
    outs = []
    for pipeline, prefetch, inp in (('2', '1', probes), ('1', '1', probes), ('1', '0', probes), ('2', '1', batches),
                                    ('3', '1', probes)):
        monkeypatch.setenv('CB_PIPELINE', pipeline)
        monkeypatch.setenv('CB_PREFETCH', prefetch)
        scf = SetCoverFilter(mismatches=3, lcf_thres=40, cover_extension=10)
        scf._ctx = ctx
        np.random.seed(5)
        pyrandom.seed(5)
        out = scf.filter(inp, genomes, input_is_grouped=True)
        outs.append(([[p.seq_str for p in g] for g in out], int(np.random.randint(0, 1 << 30))))
        assert all(s is not None for s in scf.last_stats)
    assert all(o == outs[0] for o in outs[1:])
    assert all(len(g) > 0 for g in outs[0][0])


def test_pipelining_falls_back_on_mixed_probe_lengths(ctx, monkeypatch):
    """The draw replay takes a probe list to be as uniform as its length samples say; a list with a few probes of
    another length that the samples miss makes the filter redo the call with the plain loop (which measures every
    list) from the saved RNG state: same result as CB_PIPELINE=1."""
    import random as pyrandom
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    groups = helpers.synthetic_taxa(3, 10, seed=11, length_range=(1500, 2500))
    genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
    cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
    cands[1][7] = cands[1][7][:80]                        # one shorter probe, not at a sampled position
    cands[1][-3] = cands[1][-3][:90]
    probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
    outs = []
    for pipeline in ('2', '1'):
        monkeypatch.setenv('CB_PIPELINE', pipeline)
        scf = SetCoverFilter(mismatches=3, lcf_thres=40, cover_extension=10)
        scf._ctx = ctx
        np.random.seed(6)
        pyrandom.seed(6)
        out = scf.filter(probes, genomes, input_is_grouped=True)
        outs.append(([[p.seq_str for p in g] for g in out], int(np.random.randint(0, 1 << 30))))
    assert outs[0] == outs[1]


@pytest.mark.parametrize('n_genomes', [20, 45, 100, 200, 400, 800])
def test_merge_every_list_length_class(ctx, n_genomes):
    """The warp merge sorts a probe's ranges in registers with 1, 2, 4, 8 or 16 ranges per lane, lists of 513..1024
    in a second kernel: one case per class (a probe hits every genome about once), against the oracle."""
    O = _oracle()
    rng = random.Random(100 + n_genomes)
    anc = ''.join(rng.choice('ACGT') for _ in range(400))
    genomes = [[helpers.mutate(rng, anc, 0.015)] for _ in range(n_genomes)]
    probe_strs = list(dict.fromkeys(helpers.tile_candidates([g[0] for g in genomes[:5]], 60, 30)))
    params = dict(mismatches=2, lcf_thres=45, island_of_exact_match=0, cover_extension=10, kmer_probe_map_k=15)
    np.random.seed(9)
    k, seeds, _ = O.choose_seeds(probe_strs, 2, 45, min_k=15, k=15)
    want = O.make_sets_quads(O.SeedMap(probe_strs, seeds, k), genomes, 2, 45, 0, 10)
    np.random.seed(9)
    got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
    cover.free()
    counts = np.bincount(want[:, 0], minlength=len(probe_strs))
    assert counts.max() > n_genomes // 2
    assert np.array_equal(got, want)


def test_range_list_sized_from_previous_scan_overflows_gracefully(ctx):
    """Stage A sizes its range list from the density of the previous scan with the same parameters and only counts
    when that overflows: a scan that finds almost nothing followed by a dense one (and the other way round) must give
    the same cover as a context that always counts first."""
    O = _oracle()
    rng = random.Random(77)
    anc = ''.join(rng.choice('ACGT') for _ in range(3000))
    dense = [[helpers.mutate(rng, anc, 0.01)] for _ in range(220)]
    # almost nothing to find, but not nothing (a density of zero is not used as a hint)
    sparse = [[''.join(rng.choice('ACGT') for _ in range(3000))] for _ in range(219)] + [dense[0]]
    probe_strs = list(dict.fromkeys(helpers.tile_candidates([g[0] for g in dense[:4]], 75, 25)))
    params = dict(mismatches=2, lcf_thres=60, island_of_exact_match=0, cover_extension=0, kmer_probe_map_k=20)
    np.random.seed(3)
    k, seeds, _ = O.choose_seeds(probe_strs, 2, 60, min_k=20, k=20)
    want = O.make_sets_quads(O.SeedMap(probe_strs, seeds, k), dense, 2, 60, 0, 0)
    for genomes in (sparse, dense, sparse, dense):
        np.random.seed(3)
        got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
        cover.free()
        if genomes is dense:
            assert np.array_equal(got, want)
            assert st.n_raw_ranges > 70000              # more than the hint of the sparse scan allows for (~65.6 k)
        else:
            assert 0 < st.n_raw_ranges < 2000


def test_merge_block_paths_on_tandem_repeats(ctx):
    """A probe made of a short repeat unit against tandem-repeat genomes matches on thousands of diagonals: its range
    list goes to the block-per-probe merge -- sorted in shared memory up to 12 288 ranges (the second case: ~4 600 per
    probe), in place in global memory beyond (the first case: ~14 000 per probe).  All of them collapse to one
    interval per genome."""
    O = _oracle()
    rng = random.Random(41)
    unit = 'ACGGTCA'
    flank = ''.join(rng.choice('ACGT') for _ in range(300))
    probe_strs = [(unit * 12)[:60], (unit * 12)[3:63], flank[10:70], flank[100:160]]
    params = dict(mismatches=1, lcf_thres=50, island_of_exact_match=0, cover_extension=0, kmer_probe_map_k=15)
    for scale, at_least in ((3, 2 * 12288), (1, 2 * 4000)):
        genomes = [[flank + unit * (720 * scale) + flank[::-1]], [unit * (1450 * scale) + flank],
                   [flank + unit * (2450 * scale)]]
        np.random.seed(2)
        k, seeds, _ = O.choose_seeds(probe_strs, 1, 50, min_k=15, k=15)
        want = O.make_sets_quads(O.SeedMap(probe_strs, seeds, k), genomes, 1, 50, 0, 0)
        np.random.seed(2)
        got, cover, st = _device_quads(ctx, probe_strs, genomes, params)
        cover.free()
        assert st.n_raw_ranges > at_least                      # the two repeat probes: one range per matching diagonal
        assert np.array_equal(got, want)
