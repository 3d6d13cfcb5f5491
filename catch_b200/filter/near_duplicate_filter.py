"""Near-duplicate filters on the device: drop-ins for catch/filter/near_duplicate_filter.py.

Same constructors and public attributes as the reference (`NearDuplicateFilterWithMinHash(dist_thres,
kmer_size=10)` :163, `NearDuplicateFilterWithHammingDistance(dist_thres, probe_length)` :115, `.k`,
`.reporting_prob` are read and mutated by callers).  `_filter(input)` keeps the highest-multiplicity
probe of every LSH-connected neighbourhood exactly as the reference's sequential loop does (:47-103).

Host work kept here: drawing the hash-function parameters from Python's `random` in the reference's
call order (utils/lsh.py:28,95-96,224,284-287) and rebuilding the Python-set order of the result
(:103).  Grouping identical sequences and the multiplicity ordering (:61-66) run in the library.
The MinHash inner hash is CPython's str hash; like the reference this is only reproducible under
PYTHONHASHSEED=0, which is what the device implements (SipHash-1-3 with a zero key).
"""
import math
import random

import numpy as np

from catch_b200 import _lib
from catch_b200 import coverage as cov
from catch_b200.filter.base_filter import BaseFilter
from catch_b200.probe_batch import ProbeBatch


def _num_tables(P1, k, reporting_prob):
    """utils/lsh.py:270-277."""
    if P1 == 1.0:
        return 1
    return int(math.ceil(math.log(1.0 - reporting_prob, 1.0 - math.pow(P1, k))))


def set_ordered(probes):
    """list(set built by .add() in list order) for DISTINCT probes, without a Python-level __hash__ / __eq__ call per
    element: a set of the sequences (str: hashed and compared in C) filled in the same order has the same table
    layout, hence the same iteration order, as the set of Probe objects -- a Probe hashes as its sequence
    (probe.py:324-329) and set insertion depends on nothing but the hashes and the insertion order."""
    strs = [p.seq_str for p in probes]
    by_str = dict(zip(strs, probes))
    if len(by_str) != len(strs):                      # equal sequences: keep the plain construction
        out = set()
        for p in probes:
            out.add(p)
        return list(out)
    order = set()
    order.update(strs)
    return [by_str[s] for s in order]


class NearDuplicateFilter(BaseFilter):
    def __init__(self, k, reporting_prob=0.80):
        self.k = k
        self.reporting_prob = reporting_prob
        # groupings are processed in-process (a CUDA context must not cross a fork); BaseFilter
        # keeps the reference's descending-size order
        self._ctx = None
        self.last_stats = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _draw_and_run(self, ctx, raw, off, lens):
        raise NotImplementedError

    def _filter(self, input):
        """One library call per probe list: the sequences (duplicates included, list order) are gathered
        into one buffer; grouping by sequence, the multiplicity ordering (:61-66), the LSH tables and
        the sequential keep/drop loop (:81-96) all run in cb_neardup_filter.  It returns, in priority
        order, the list index of the first occurrence of every kept sequence -- the Probe object the
        reference's dict keyed by Probe keeps."""
        if not isinstance(input, ProbeBatch):
            input = list(input)
        if not len(input):
            # the reference still builds the lookup (and draws its parameters) for empty input
            self._draw_only()
            return []
        ctx = self._context()
        gathered = cov.gather_staged(ctx, 0, input)
        if gathered is None:
            gathered = cov.gather_probes(input)
        raw, lens = gathered[0], gathered[1]
        off = cov.offsets_from_lengths(lens)
        kept, n_distinct, st = self._draw_and_run(ctx, raw, off, lens)
        self.last_stats = st.as_dict()
        self.last_stats['n_distinct'] = n_distinct
        if isinstance(input, ProbeBatch):
            # same order as the reference's list(set(...)): the set order of the kept SEQUENCES (see set_ordered)
            strs = input.strs(kept)
            rank = {s: i for i, s in enumerate(strs)}
            order = set()
            order.update(strs)
            return input.take(kept[np.fromiter((rank[s] for s in order), dtype=np.int64, count=len(strs))])
        return set_ordered([input[i] for i in kept.tolist()])     # :96-103: to_include.add(p) ..., list(to_include)


class NearDuplicateFilterWithHammingDistance(NearDuplicateFilter):
    def __init__(self, dist_thres, probe_length):
        super().__init__(k=20)
        self.probe_length = probe_length
        self.dist_thres = dist_thres

    def _params(self):
        P1 = 1.0 - float(self.dist_thres) / float(self.probe_length)      # utils/lsh.py:45
        n_tab = _num_tables(P1, self.k, self.reporting_prob)
        pos = [random.randint(0, self.probe_length - 1) for _ in range(n_tab * self.k)]   # :28
        return n_tab, np.array(pos, dtype=np.int32)

    def _draw_only(self):
        self._params()

    def _draw_and_run(self, ctx, raw, off, lens):
        n_tab, pos = self._params()
        if np.any(lens != self.probe_length):
            raise AssertionError("all probes must have length %d" % self.probe_length)    # utils/lsh.py:30
        return ctx.neardup_filter(raw, off, 1, None, None, pos, n_tab, self.k, 0, self.dist_thres)


class NearDuplicateFilterWithMinHash(NearDuplicateFilter):
    def __init__(self, dist_thres, kmer_size=10):
        super().__init__(k=3)
        self.kmer_size = kmer_size
        self.dist_thres = dist_thres

    def _params(self):
        n_tab = _num_tables(1.0 - self.dist_thres, self.k, self.reporting_prob)           # utils/lsh.py:164
        p = 2 ** 31 - 1
        a, b = [], []
        for _ in range(n_tab * self.k):
            a.append(random.randint(1, p))                                                # :95
            b.append(random.randint(0, p))                                                # :96
        return n_tab, np.array(a, dtype=np.uint32), np.array(b, dtype=np.uint32)

    def _draw_only(self):
        self._params()

    def _draw_and_run(self, ctx, raw, off, lens):
        n_tab, a, b = self._params()
        if np.any(lens < self.kmer_size):
            raise AssertionError("kmer_size exceeds the length of a probe")               # utils/lsh.py:117
        return ctx.neardup_filter(raw, off, 0, a, b, None, n_tab, self.k, self.kmer_size, self.dist_thres)
