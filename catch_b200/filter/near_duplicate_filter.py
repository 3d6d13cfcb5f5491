"""Near-duplicate filters on the device: drop-ins for catch/filter/near_duplicate_filter.py.

Same constructors and public attributes as the reference (`NearDuplicateFilterWithMinHash(dist_thres,
kmer_size=10)` :163, `NearDuplicateFilterWithHammingDistance(dist_thres, probe_length)` :115, `.k`,
`.reporting_prob` are read and mutated by callers).  `_filter(input)` keeps the highest-multiplicity
probe of every LSH-connected neighbourhood exactly as the reference's sequential loop does (:47-103).

Host work kept here: multiplicity ordering (:61-66), drawing the hash-function parameters from
Python's `random` in the reference's call order (utils/lsh.py:28,95-96,224,284-287) and rebuilding
the Python-set order of the result (:103).  The MinHash inner hash is CPython's str hash; like the
reference this is only reproducible under PYTHONHASHSEED=0, which is what the device implements
(SipHash-1-3 with a zero key).
"""
import collections
import math
import operator
import random

import numpy as np

from catch_b200 import _lib
from catch_b200 import coverage as cov
from catch_b200.filter.base_filter import BaseFilter


def _num_tables(P1, k, reporting_prob):
    """utils/lsh.py:270-277."""
    if P1 == 1.0:
        return 1
    return int(math.ceil(math.log(1.0 - reporting_prob, 1.0 - math.pow(P1, k))))


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

    def _draw_and_run(self, ctx, buf, off):
        raise NotImplementedError

    def _filter(self, input):
        # multiplicity, descending, stable in first-occurrence order (:61-66).  Counting runs on the
        # sequence strings (C speed); the objects returned are the first-seen Probe of each sequence,
        # as with the reference's dict keyed by Probe.
        input = list(input)
        strs = [p.seq_str for p in input]
        occurrences = collections.Counter(strs)
        order = [s for s, _ in sorted(occurrences.items(), key=operator.itemgetter(1), reverse=True)]
        if not order:
            # the reference still builds the lookup (and draws its parameters) for empty input
            self._draw_only()
            return []
        first_seen = dict(zip(reversed(strs), reversed(input)))
        buf, off, _ = cov._concat_ascii(order)
        keep, st = self._draw_and_run(self._context(), buf, off)
        self.last_stats = st.as_dict()
        to_include = set()
        for s, k in zip(order, keep.tolist()):
            if k:
                to_include.add(first_seen[s])
        return list(to_include)                                   # :103


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

    def _draw_and_run(self, ctx, buf, off):
        n_tab, pos = self._params()
        if np.any(np.diff(off) != self.probe_length):
            raise AssertionError("all probes must have length %d" % self.probe_length)    # utils/lsh.py:30
        return ctx.hamming_neardup(buf, off, pos, n_tab, self.k, self.dist_thres)


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

    def _draw_and_run(self, ctx, buf, off):
        n_tab, a, b = self._params()
        if np.any(np.diff(off) < self.kmer_size):
            raise AssertionError("kmer_size exceeds the length of a probe")               # utils/lsh.py:117
        return ctx.minhash_neardup(buf, off, a, b, n_tab, self.k, self.kmer_size, self.dist_thres)
