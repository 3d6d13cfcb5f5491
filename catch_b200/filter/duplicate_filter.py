"""Exact-duplicate removal on the device: drop-in for catch/filter/duplicate_filter.py:16-26.

The reference returns list(OrderedDict.fromkeys(input)): the first occurrence of every distinct
probe, in input order (Probe hashes and compares by sequence, probe.py:324-329).  Here the
sequences are gathered into one buffer and grouped by cb_group_duplicates (exact comparison of the
packed bit planes in a device hash table); the Probe objects returned are the first occurrences.
"""
import numpy as np

from catch_b200 import _lib
from catch_b200 import coverage as cov
from catch_b200.filter.base_filter import BaseFilter
from catch_b200.probe_batch import ProbeBatch


class DuplicateFilter(BaseFilter):
    def __init__(self):
        self._ctx = None
        self.last_stats = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _filter(self, input):
        if not isinstance(input, (list, tuple, ProbeBatch)):
            input = list(input)
        if not len(input):
            return []
        ctx = self._context()
        gathered = cov.gather_staged(ctx, 0, input)
        if gathered is None:
            gathered = cov.gather_probes(input)
        off = cov.offsets_from_lengths(gathered[1])
        first, _, st = ctx.group_duplicates(gathered[0], off)
        self.last_stats = st.as_dict()
        if isinstance(input, ProbeBatch):
            return input.take(first)                # still one buffer; no Probe objects
        return [input[i] for i in first.tolist()]
