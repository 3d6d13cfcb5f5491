"""Greedy (multi-universe) set cover on the device: drop-ins for catch/utils/set_cover.py.

`approx_multiuniverse(sets, costs, universe_p, ranks, use_arrays, use_intervalsets)` (:147-615) and
`approx(sets, costs, p)` (:18-145) keep the reference's arguments, validation errors and return
value (a Python set of the chosen set identifiers, built by .add() in pick order).  The elements
are turned into integer intervals per universe, handed to cb_cover_import, and the greedy loop runs
in cb_setcover / cb_setcover_costs (csrc/setcover.cu).  There is no CPU fallback.

Tie-breaking follows the reference exactly: it scans `set(sets.keys())` with a strict '<', so among
equal ratios the identifier that comes first in that Python set's iteration order wins; the device
breaks ties by smallest index, and indices are assigned in that iteration order.
"""
import numpy as np

from catch_b200 import _lib


def _as_intervals(s, use_intervalsets, index_of):
    """One set's elements in one universe -> list of (start, end) integer intervals."""
    if use_intervalsets:
        if isinstance(s, tuple):
            return [(int(s[0]), int(s[1]))]                      # a single interval (:405-407)
        return [(int(a), int(b)) for a, b in s.intervals]        # IntervalSet
    idx = np.unique(np.fromiter((index_of[v] for v in s), dtype=np.int64))
    if idx.size == 0:
        return []
    cut = np.flatnonzero(np.diff(idx) != 1) + 1                  # runs of consecutive indices
    starts = np.concatenate(([0], cut))
    ends = np.concatenate((cut, [idx.size]))
    return [(int(idx[a]), int(idx[b - 1]) + 1) for a, b in zip(starts, ends)]


def approx_multiuniverse(sets, costs=None, universe_p=None, ranks=None, use_arrays=False,
                         use_intervalsets=False, logger_prefix="", ctx=None):
    if use_arrays and use_intervalsets:
        raise ValueError("Cannot use both arrays and IntervalSets")
    if costs is not None:
        for c in costs.values():
            if c < 0:
                raise ValueError("All costs must be nonnegative")
        for set_id in sets.keys():
            if set_id not in costs:
                raise ValueError("costs is missing a value for set %d" % set_id)
    # universes and, without interval sets, a dense integer index for the elements of each
    universe_ids, index_of = [], {}
    seen_u = {}
    for by_u in sets.values():
        for u, s in by_u.items():
            if u not in seen_u:
                seen_u[u] = len(universe_ids)
                universe_ids.append(u)
                index_of[u] = {}
            if not use_intervalsets:
                m = index_of[u]
                for v in s:
                    if v not in m:
                        m[v] = len(m)
    if universe_p is not None:
        for p in universe_p.values():
            if p < 0 or p > 1:
                raise ValueError("The coverage fraction (p) of each universe must be in [0,1]")
        for u in universe_ids:
            if u not in universe_p:
                raise ValueError("universe_p is missing a value for universe %d" % u)
    if ranks is not None:
        for set_id in sets.keys():
            if set_id not in ranks:
                raise ValueError("ranks is missing a value for set %d" % set_id)
    order = list(set(sets.keys()))                 # the reference's scan order (:483), ties go to the first
    if not order or not universe_ids:
        return set()
    pid, gen, start, end = [], [], [], []
    glen = np.zeros(len(universe_ids), dtype=np.int64)
    for i, set_id in enumerate(order):
        for u, s in sets[set_id].items():
            j = seen_u[u]
            for a, b in _as_intervals(s, use_intervalsets, index_of[u]):
                if a < 0:
                    raise ValueError("interval coordinates must be non-negative")
                if b > a:
                    pid.append(i)
                    gen.append(j)
                    start.append(a)
                    end.append(b)
                    if b > glen[j]:
                        glen[j] = b
    ctx = ctx if ctx is not None else _lib.default_context()
    cover = ctx.cover_import(len(order), glen, pid, gen, start, end)
    try:
        c = None if costs is None else np.array([float(costs[s]) for s in order], dtype=np.float64)
        r = None if ranks is None else np.array([ranks[s] for s in order], dtype=np.int32)
        up = None if universe_p is None else np.array([float(universe_p[u]) for u in universe_ids], dtype=np.float64)
        picks, _ = ctx.setcover(cover, len(order), r, up, costs=c)
    finally:
        cover.free()
    chosen = set()
    for i in picks.tolist():
        chosen.add(order[i])
    return chosen


def approx(sets, costs=None, p=1.0, ctx=None):
    """Single-universe weighted partial set cover (catch/utils/set_cover.py:18-145): the same greedy
    rule with one universe."""
    if p < 0 or p > 1:
        raise ValueError("p must be in [0,1]")
    if costs is not None:
        for c in costs.values():
            if c < 0:
                raise ValueError("All costs must be nonnegative")
    return approx_multiuniverse({k: {0: s} for k, s in sets.items()}, costs=costs, universe_p={0: p}, ctx=ctx)
