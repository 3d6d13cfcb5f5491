"""Loading of the fixtures written by tests/golden/make_golden.py (outputs of the reference)."""
import gzip
import json
import os

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    with gzip.open(os.path.join(HERE, name), 'rt') as f:
        return json.load(f)


def setcover_case_to_quads(rec):
    """A recorded approx_multiuniverse call -> (quads, n_sets, n_universes, costs, universe_p, ranks,
    set id list).  Set ids are remapped to 0..n-1 in ascending order (the reference iterates a set
    of small ints ascending, utils/set_cover.py:483)."""
    set_ids = sorted(int(s) for s in rec['sets'])
    sid = {s: i for i, s in enumerate(set_ids)}
    n_u = 0
    quads = []
    for s, by_u in rec['sets'].items():
        for u, ivs in by_u.items():
            n_u = max(n_u, int(u) + 1)
            for a, b in ivs:
                quads.append((sid[int(s)], int(u), int(a), int(b)))
    if rec['universe_p'] is not None:
        n_u = max([n_u] + [int(u) + 1 for u in rec['universe_p']])
    costs = None if rec['costs'] is None else [rec['costs'][str(s)] for s in set_ids]
    ranks = None if rec['ranks'] is None else [rec['ranks'][str(s)] for s in set_ids]
    up = None
    if rec['universe_p'] is not None:
        up = [rec['universe_p'].get(str(u), 1.0) for u in range(n_u)]
    return quads, len(set_ids), n_u, costs, up, ranks, set_ids
