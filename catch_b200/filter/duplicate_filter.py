"""Exact-duplicate removal (reference: catch/filter/duplicate_filter.py:20-26); host op."""
from collections import OrderedDict

from catch_b200.filter.base_filter import BaseFilter


class DuplicateFilter(BaseFilter):
    def _filter(self, input):
        return list(OrderedDict.fromkeys(input))
