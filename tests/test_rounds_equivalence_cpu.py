"""The parallel-rounds formulation of the greedy set cover (csrc/setcover.cu:
greedy_par_kernel) restated in plain Python and checked against the sequential rule of
utils/set_cover.py:393-526 (max gain, smallest id) on random interval instances: same picks, and
sorting the picks by their pick-time key restores the sequential ORDER.  No GPU, no library: this
pins the algorithmic claim itself -- candidate list = prefix of the key order, conflicts detected at
64-bit word granularity, every local maximum of a round applied at once."""
import random

import numpy as np


def _instance(rng, n_sets, n_universes, length, max_iv):
    sets = []
    for _ in range(n_sets):
        ivs = []
        for _ in range(rng.randint(1, max_iv)):
            u = rng.randrange(n_universes)
            a = rng.randrange(length - 1)
            b = min(length, a + rng.randint(1, 40))
            ivs.append((u * (length + 64), a, b))               # universes never share a word
        # merge overlapping intervals of the set (the cover holds disjoint intervals per set)
        ivs.sort()
        merged = []
        for base, a, b in ivs:
            if merged and merged[-1][0] == base and a <= merged[-1][2]:
                merged[-1][2] = max(merged[-1][2], b)
            else:
                merged.append([base, a, b])
        sets.append([(base + a, base + b) for base, a, b in merged])
    return sets


def _sequential(sets, size):
    U = np.zeros(size, dtype=bool)
    for ivs in sets:
        for a, b in ivs:
            U[a:b] = True
    picks = []
    while U.any():
        gains = [sum(int(U[a:b].sum()) for a, b in ivs) for ivs in sets]
        w = max(range(len(sets)), key=lambda i: (gains[i], -i))
        assert gains[w] > 0
        picks.append(w)
        for a, b in sets[w]:
            U[a:b] = False
    return picks


def _rounds(sets, size, list_cap):
    U = np.zeros(size, dtype=bool)
    for ivs in sets:
        for a, b in ivs:
            U[a:b] = True
    gain = lambda i: sum(int(U[a:b].sum()) for a, b in sets[i])     # noqa: E731
    keyed, n_rounds = [], 0
    while U.any():
        gains = [gain(i) for i in range(len(sets))]
        gmax = max(gains)
        # candidate list: everything with gain >= tau, tau chosen so that the list fits (a prefix
        # of the key order); with more ties at gmax than fit, the lowest ids among them
        order = sorted((i for i in range(len(sets)) if gains[i] > 0), key=lambda i: (-gains[i], i))
        tau = gmax
        for t in sorted(set(gains), reverse=True):
            if t > 0 and sum(g >= t for g in gains) <= list_cap:
                tau = t
        cand = [i for i in order if gains[i] >= tau][:list_cap]
        while True:                                                 # rounds on this list
            active = [i for i in cand if gain(i) >= tau]
            if not active:
                break
            n_rounds += 1
            key = {i: (gain(i), -i) for i in active}
            mark = {}
            for i in active:                                        # mark: max key per 64-bit word with uncovered bits
                for a, b in sets[i]:
                    for w in range(a >> 6, ((b - 1) >> 6) + 1):
                        lo, hi = max(a, w << 6), min(b, (w + 1) << 6)
                        if U[lo:hi].any() and mark.get(w, (0, 0)) < key[i]:
                            mark[w] = key[i]
            winners = []
            for i in active:                                        # check
                ok = True
                for a, b in sets[i]:
                    for w in range(a >> 6, ((b - 1) >> 6) + 1):
                        lo, hi = max(a, w << 6), min(b, (w + 1) << 6)
                        if U[lo:hi].any() and mark[w] != key[i]:
                            ok = False
                if ok:
                    winners.append(i)
            assert winners                                          # the largest key always wins
            for i in winners:                                       # apply all at once
                keyed.append((key[i], i))
                for a, b in sets[i]:
                    U[a:b] = False
    keyed.sort(key=lambda t: t[0], reverse=True)
    return [i for _, i in keyed], n_rounds


def test_parallel_rounds_equal_sequential_greedy():
    rng = random.Random(2024)
    total_picks = total_rounds = 0
    for case in range(40):
        n_universes = rng.choice([1, 2, 5])
        length = rng.choice([100, 300, 700])
        sets = _instance(rng, rng.choice([5, 30, 80]), n_universes, length, rng.choice([1, 3, 6]))
        size = n_universes * (length + 64)
        want = _sequential(sets, size)
        for cap in (1, 4, 1000):
            got, n_rounds = _rounds(sets, size, cap)
            assert got == want, (case, cap)
        total_picks += len(want)
        total_rounds += n_rounds
    assert total_rounds < total_picks                               # rounds really do batch picks
