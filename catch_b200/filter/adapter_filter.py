"""AdapterFilter on the device scan: the drop-in for catch/filter/adapter_filter.py:120-393.

Same constructor and `_filter(input, target_genomes)` contract: every probe comes back with the 'A'
or the 'B' adapter on both ends, chosen by votes over all target sequences (interval scheduling per
sequence, flip of a sequence's votes when that makes the plurality clearer; see the reference's
module docstring for the rationale).  The probe-vs-sequence scan -- where the reference spends its
time (probe.find_probe_covers_in_sequence per sequence, adapter_filter.py:214) -- runs on the GPU
(cb_coverage_records); the scheduling and the vote bookkeeping are cheap and stay on the host,
vectorised per sequence.

Tie-breaking is the reference's: interval.schedule sorts by end point with a STABLE sort
(utils/interval.py:336), so intervals with equal ends keep the order in which they were listed,
which is the order in which find_probe_covers_in_sequence first saw each probe (its result dict is
filled while scanning the sequence from left to right, probe.py:1060-1106, chunks of the worker pool
concatenated in position order, :1240-1248): by position of the first successful seed hit, then by
the order of the probes in the k-mer map, i.e. in the input list.
"""
import logging

import numpy as np

from catch_b200 import _lib
from catch_b200 import coverage as cov
from catch_b200.filter.base_filter import BaseFilter

logger = logging.getLogger(__name__)


def merged_ranges_per_probe(rec):
    """rec: records of ONE sequence, int64 [n, 5].  Returns (probe, start, end, first_hit) arrays of the
    merged cover ranges (overlapping or touching ranges united, utils/interval.py:304-314), sorted by
    (probe, start); first_hit[i] is the smallest hit position of the range's probe in this sequence."""
    if len(rec) == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z, z
    order = np.lexsort((rec[:, 3], rec[:, 2], rec[:, 0]))
    p, s, e, h = rec[order, 0], rec[order, 2], rec[order, 3], rec[order, 4]
    new_probe = np.r_[True, p[1:] != p[:-1]]
    # running maximum of the ends inside each probe's block (ends are < 2^32: offset the blocks apart)
    block = np.cumsum(new_probe) - 1
    shifted = e + block * (1 << 33)
    run_max = np.maximum.accumulate(shifted) - block * (1 << 33)
    starts_group = np.r_[True, (s[1:] > run_max[:-1]) | new_probe[1:]]
    gid = np.cumsum(starts_group) - 1
    n_groups = int(gid[-1]) + 1
    g_start = s[starts_group]
    g_probe = p[starts_group]
    g_end = np.zeros(n_groups, dtype=np.int64)
    np.maximum.at(g_end, gid, e)
    first_hit_of_probe = np.full(int(p.max()) + 1, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first_hit_of_probe, p, h)
    return g_probe, g_start, g_end, first_hit_of_probe[g_probe]


def schedule(starts, ends):
    """Greedy interval scheduling (utils/interval.py:319-358) on intervals ALREADY in the order
    interval.schedule's stable sort would leave them (ascending end, ties in listing order): indices of the
    chosen intervals."""
    n = len(starts)
    chosen = []
    i = 0
    last_end = None
    while i < n:
        if last_end is None:
            chosen.append(i)
            last_end = ends[i]
            i += 1
            continue
        ok = starts[i:] >= last_end
        if not ok.any():
            break
        i += int(np.argmax(ok))
        chosen.append(i)
        last_end = ends[i]
        i += 1
    return chosen


class AdapterFilter(BaseFilter):
    def __init__(self, adapter_a, adapter_b, mismatches, lcf_thres, island_of_exact_match=0,
                 custom_cover_range_fn=None, kmer_probe_map_k=20):
        if len(adapter_a) != 2 or len(adapter_b) != 2:
            raise ValueError(("adapter_a/adapter_b arguments must be tuples "
                              "of length 2, giving the sequences to add onto "
                              "the 5' and 3' ends"))
        if custom_cover_range_fn is not None:
            raise NotImplementedError("custom hybridization functions are Python callables and "
                                      "are not supported by the device implementation")
        self.adapter_a_5end, self.adapter_a_3end = adapter_a
        self.adapter_b_5end, self.adapter_b_3end = adapter_b
        self.mismatches = mismatches
        self.lcf_thres = lcf_thres
        self.island_of_exact_match = island_of_exact_match
        self.kmer_probe_map_k = kmer_probe_map_k
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _votes_in_sequence(self, n_probes, rep, rec, sequence=None, kmer_order=None):
        """adapter_filter.py:191-238 for one sequence: int64 [n_probes, 2] of (A, B) votes."""
        votes = np.zeros((n_probes, 2), dtype=np.int64)
        g_probe, g_start, g_end, g_hit = merged_ranges_per_probe(rec)
        if len(g_probe) == 0:
            return votes
        # listing order of the reference: probes by first successful hit, probes first seen at the same
        # position in the order of the k-mer map's set (coverage.KmerMapOrder), ranges ascending; then the
        # stable sort by end
        g_tie = cov.listing_tie_ranks(g_probe, g_hit, sequence, kmer_order)
        listing = np.lexsort((g_start, g_probe, g_tie, g_hit))
        by_end = listing[np.argsort(g_end[listing], kind='stable')]
        chosen = schedule(g_start[by_end], g_end[by_end])
        chosen_probes = np.unique(g_probe[by_end][chosen])
        aligned = np.unique(g_probe)
        a_vote = np.zeros(n_probes, dtype=bool)
        b_vote = np.zeros(n_probes, dtype=bool)
        a_vote[chosen_probes] = True
        b_vote[aligned] = True
        b_vote &= ~a_vote
        if rep is not None:                     # probes with the same sequence are one key of the reference's dicts
            a_vote, b_vote = a_vote[rep], b_vote[rep]
        votes[:, 0] = a_vote
        votes[:, 1] = b_vote
        return votes

    def _make_votes_across_target_genomes(self, probes, target_genomes):
        """adapter_filter.py:299-362.  Returns a list of (A, B) tuples, one per probe."""
        ctx = self._context()
        probe_strs = [p.seq_str for p in probes]
        n = len(probe_strs)
        cumulative = np.zeros((n, 2), dtype=np.int64)
        if n == 0:
            return []
        plan = cov.SeedPlan(probe_strs, self.mismatches, self.lcf_thres, self.kmer_probe_map_k)
        rep = np.asarray(plan.rep) if plan.rep is not None else None
        kmer_order = cov.KmerMapOrder(probes, plan)
        seqs = [seq for genomes_from_group in target_genomes for g in genomes_from_group for seq in g.seqs]
        for batch_start, batch in _batches_with_offsets(seqs):
            rec = cov.scan_records(ctx, probe_strs, batch, plan, self.mismatches, self.lcf_thres,
                                   self.island_of_exact_match)
            order = np.argsort(rec[:, 1], kind='stable')
            rec = rec[order]
            bounds = np.searchsorted(rec[:, 1], np.arange(len(batch) + 1))
            for q in range(len(batch)):
                votes = self._votes_in_sequence(n, rep, rec[bounds[q]:bounds[q + 1]], batch[q], kmer_order)
                flipped = votes[:, ::-1]
                with_nonflipped = cumulative + votes
                with_flipped = cumulative + flipped
                if with_flipped.max(axis=1).sum() > with_nonflipped.max(axis=1).sum():
                    cumulative = with_flipped
                else:
                    cumulative = with_nonflipped
        return [tuple(int(x) for x in row) for row in cumulative]

    def _filter(self, input, target_genomes):
        input = list(input)
        logger.info("Computing adapter votes across all target genomes")
        votes = self._make_votes_across_target_genomes(input, target_genomes)
        logger.info("Adding adapters to probes based on votes")
        out = []
        for p, vote in zip(input, votes):
            if vote[0] > vote[1]:
                out.append(p.with_prepended_str(self.adapter_a_5end).with_appended_str(self.adapter_a_3end))
            else:
                out.append(p.with_prepended_str(self.adapter_b_5end).with_appended_str(self.adapter_b_3end))
        return out


def _batches_with_offsets(seqs, max_bases=1 << 28):
    start = 0
    for batch in cov.sequence_batches(seqs, max_bases):
        yield start, batch
        start += len(batch)
