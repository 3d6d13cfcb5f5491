"""CPU oracle for the CATCH probe-coverage + set-cover + near-duplicate hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
__graft_entry__.py and the cpu_baseline / `--impl reference` legs of bench.py may
import it; catch_b200/ never does.

The heavy loops live in catch_oracle.c (plain C, built by oracle/Makefile); the thin
Python here restates the host-side bookkeeping of the reference (RNG draws, duplicate
handling, Python-set ordering) so results can be compared object-for-object with the
reference.  Parity status: PINNED against the reference's own tests' golden vectors
and against outputs of the reference itself (tests/golden/, tests/test_oracle_*.py).

All file:line citations are into the reference tree, /root/reference/catch/...
"""
import ctypes
import math
import os
import pickle
import random
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'catch_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'liboracle.so'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        c = ctypes
        L.orc_k_lcf_around_anchor.argtypes = [c.c_char_p, c.c_int, c.c_char_p, c.c_int, c.c_int,
                                              c.c_int, c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_int)]
        L.orc_lcf_cover.argtypes = [c.c_char_p, c.c_int, c.c_char_p, c.c_int, c.c_int, c.c_int, c.c_int,
                                    c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_int),
                                    c.POINTER(c.c_int)]
        L.orc_seedmap_build.restype = c.c_void_p
        L.orc_seedmap_build.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_void_p, c.c_void_p]
        L.orc_seedmap_free.argtypes = [c.c_void_p]
        L.orc_find_probe_covers_in_sequence.argtypes = [
            c.c_void_p, c.c_char_p, c.c_int64, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int,
            c.POINTER(c.POINTER(c.c_int64)), c.POINTER(c.c_int64)]
        L.orc_make_sets.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p, c.c_int,
                                    c.c_int, c.c_int, c.c_int, c.c_int,
                                    c.POINTER(c.POINTER(c.c_int64)), c.POINTER(c.c_int64)]
        L.orc_set_cover.argtypes = [c.c_void_p, c.c_int64, c.c_int64, c.c_int64, c.c_void_p, c.c_void_p,
                                    c.c_void_p, c.c_void_p, c.POINTER(c.c_int64)]
        L.orc_free.argtypes = [c.c_void_p]
        L.orc_siphash13_zero_key.restype = c.c_uint64
        L.orc_siphash13_zero_key.argtypes = [c.c_char_p, c.c_int64]
        L.orc_abs_pyhash.restype = c.c_uint64
        L.orc_abs_pyhash.argtypes = [c.c_char_p, c.c_int64]
        L.orc_minhash.restype = c.c_uint32
        L.orc_minhash.argtypes = [c.c_char_p, c.c_int64, c.c_int, c.c_uint64, c.c_uint64]
        L.orc_near_duplicate.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_int,
                                         c.c_int, c.c_void_p, c.c_void_p, c.c_double, c.c_void_p]
        L.orc_num_threads.restype = c.c_int
        _LIB = L
    return _LIB


def num_threads():
    return lib().orc_num_threads()


def _b(s):
    return s if isinstance(s, bytes) else s.encode('latin-1')


# --------------------------------------------------------------------------- a7
def k_lcf_around_anchor(a, b, anchor_start, anchor_end, k):
    """utils/longest_common_substring.py:59-159."""
    ol, os_ = ctypes.c_int(), ctypes.c_int()
    r = lib().orc_k_lcf_around_anchor(_b(a), len(a), _b(b), len(b), anchor_start, anchor_end, k,
                                      ctypes.byref(ol), ctypes.byref(os_))
    if r != 0:
        raise ValueError("anchors are different in a and b")
    return ol.value, os_.value


# --------------------------------------------------------------------------- a6
def lcf_cover(probe_seq, sequence, kmer_start, kmer_end, full_probe_len, full_sequence_len,
              mismatches, lcf_thres, island_of_exact_match=0):
    """probe.py:1328-1344 (the function returned by
    probe_covers_sequence_by_longest_common_substring)."""
    s, e = ctypes.c_int(), ctypes.c_int()
    r = lib().orc_lcf_cover(_b(probe_seq), len(probe_seq), _b(sequence), len(sequence), kmer_start,
                            kmer_end, full_probe_len, full_sequence_len, mismatches, lcf_thres,
                            island_of_exact_match, ctypes.byref(s), ctypes.byref(e))
    if r < 0:
        raise ValueError("anchors are different in a and b")
    return (s.value, e.value) if r == 1 else None


# --------------------------------------------------------------------------- a2
class PigeonholeRequiresTooSmallKmerSizeError(Exception):
    pass


def pigeonhole_k(probe_length, mismatches, min_k):
    """k selection of probe.py:473-491."""
    if mismatches == 0:
        k = probe_length
    else:
        k = int(probe_length / mismatches)
        if k == float(probe_length) / mismatches:
            k -= 1
        while probe_length % k != 0:
            k -= 1
    if k < min_k:
        raise PigeonholeRequiresTooSmallKmerSizeError()
    return k


def choose_seeds(probe_strs, mismatches, lcf_thres, min_k=20, k=20, num_kmers_per_probe=20):
    """Seed positions per probe, probe.py:507-577 (+ :356-405, :414-504).

    Returns (k, [list of drawn positions per probe, in list order], mode).  In random mode
    this consumes numpy's legacy global stream exactly as the reference does: one
    np.random.choice(L-k+1, size=20, replace=True) per probe, in list order (:393-396).
    """
    if len(probe_strs) == 0:
        return k, [], 'empty'
    L0 = len(probe_strs[0])
    differ = any(len(p) != L0 for p in probe_strs)
    use_random = (mismatches is None or lcf_thres is None or differ or lcf_thres < L0)
    if not use_random:
        try:
            kk = pigeonhole_k(L0, mismatches, min_k)
            return kk, [list(range(0, L0, kk)) for _ in probe_strs], 'pigeonhole'
        except PigeonholeRequiresTooSmallKmerSizeError:
            pass
    seeds = []
    for p in probe_strs:
        if k > len(p):
            raise ValueError("k is larger than the length of a probe")
        n = len(p) - k + 1
        seeds.append([int(x) for x in np.random.choice(n, size=num_kmers_per_probe, replace=True)])
    return k, seeds, 'random'


def _flatten_strs(strs):
    off = np.zeros(len(strs) + 1, dtype=np.int64)
    if len(strs):
        off[1:] = np.cumsum([len(s) for s in strs])
    buf = np.frombuffer(b''.join(_b(s) for s in strs), dtype=np.uint8).copy() if len(strs) else \
        np.zeros(0, dtype=np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, dtype=np.uint8)
    return buf, off


class SeedMap:
    """probe.py:580-763 SharedKmerProbeMap over UNIQUE probe sequences.

    `probe_strs` may contain duplicates; as in the reference the map is keyed by sequence
    (Probe.__hash__/__eq__, probe.py:324-329), so the (probe, pos) sets of duplicates are
    unioned, and `rep[i]` gives for list index i the LAST list index with the same sequence
    (filter/set_cover_filter.py:408-412 `probe_id[p] = id`)."""

    def __init__(self, probe_strs, seeds, k):
        self.k = k
        last = {}
        for i, s in enumerate(probe_strs):
            last[s] = i
        self.rep = [last[s] for s in probe_strs]
        uniq = {}
        for i, s in enumerate(probe_strs):
            uniq.setdefault(s, set()).update(seeds[i])
        self.uniq_strs = list(uniq.keys())
        self.uniq_to_list_id = [last[s] for s in self.uniq_strs]
        self._buf, self._off = _flatten_strs(self.uniq_strs)
        seed_lists = [sorted(uniq[s]) for s in self.uniq_strs]
        self._seed_off = np.zeros(len(seed_lists) + 1, dtype=np.int64)
        if seed_lists:
            self._seed_off[1:] = np.cumsum([len(x) for x in seed_lists])
        flat = [x for sl in seed_lists for x in sl]
        self._seed_pos = np.array(flat if flat else [0], dtype=np.int32)
        self._h = lib().orc_seedmap_build(self._buf.ctypes.data, self._off.ctypes.data,
                                          len(self.uniq_strs), k, self._seed_off.ctypes.data,
                                          self._seed_pos.ctypes.data)
        if not self._h:
            raise MemoryError()

    def __del__(self):
        if getattr(self, '_h', None):
            lib().orc_seedmap_free(self._h)
            self._h = None


def find_probe_covers_in_sequence(seedmap, sequence, mismatches, lcf_thres, island=0,
                                  merge_overlapping=True, n_threads=1):
    """probe.py:1122-1271.  Returns {list_id_of_probe: [(start, end), ...]}."""
    out = ctypes.POINTER(ctypes.c_int64)()
    n = ctypes.c_int64()
    r = lib().orc_find_probe_covers_in_sequence(seedmap._h, _b(sequence), len(sequence), mismatches,
                                                lcf_thres, island, 1 if merge_overlapping else 0,
                                                n_threads, ctypes.byref(out), ctypes.byref(n))
    if r != 0:
        raise RuntimeError("oracle scan failed: %d" % r)
    res = {}
    if n.value:
        arr = np.ctypeslib.as_array(out, shape=(n.value * 3,)).reshape(-1, 3).copy()
        lib().orc_free(out)
        for p, s, e in arr:
            res.setdefault(seedmap.uniq_to_list_id[int(p)], []).append((int(s), int(e)))
    return res


def make_sets_quads(seedmap, genomes, mismatches, lcf_thres, island, cover_extension, n_threads=1):
    """filter/set_cover_filter.py:359-470 as a flat int64 array of
    (set_id, universe, start, end), sorted, set ids being probe LIST indices."""
    seqs, seq_genome = [], []
    for j, g in enumerate(genomes):
        for s in g:
            seqs.append(s)
            seq_genome.append(j)
    buf, off = _flatten_strs(seqs)
    sg = np.array(seq_genome if seq_genome else [0], dtype=np.int32)
    out = ctypes.POINTER(ctypes.c_int64)()
    n = ctypes.c_int64()
    r = lib().orc_make_sets(seedmap._h, buf.ctypes.data, off.ctypes.data, len(seqs), sg.ctypes.data,
                            mismatches, lcf_thres, island, cover_extension, n_threads,
                            ctypes.byref(out), ctypes.byref(n))
    if r != 0:
        raise RuntimeError("oracle make_sets failed: %d" % r)
    if n.value == 0:
        return np.zeros((0, 4), dtype=np.int64)
    arr = np.ctypeslib.as_array(out, shape=(n.value * 4,)).reshape(-1, 4).copy()
    lib().orc_free(out)
    ids = np.array(seedmap.uniq_to_list_id, dtype=np.int64)
    arr[:, 0] = ids[arr[:, 0]]
    order = np.lexsort((arr[:, 2], arr[:, 1], arr[:, 0]))
    return arr[order]


def quads_to_sets(quads, n_sets):
    """The dict-of-dicts shape _make_sets returns (single interval kept as a tuple)."""
    sets = {i: {} for i in range(n_sets)}
    for s, u, a, b in quads:
        sets[int(s)].setdefault(int(u), []).append((int(a), int(b)))
    for s in sets:
        for u in sets[s]:
            if len(sets[s][u]) == 1:
                sets[s][u] = sets[s][u][0]
    return sets


# --------------------------------------------------------------------------- a11
def set_cover_quads(quads, n_sets, n_universes, costs=None, universe_p=None, ranks=None):
    """utils/set_cover.py:147-615 approx_multiuniverse(use_intervalsets=True); returns the
    chosen set ids in pick order."""
    q = np.ascontiguousarray(quads, dtype=np.int64).reshape(-1, 4)
    order = np.lexsort((q[:, 2], q[:, 1], q[:, 0]))
    q = np.ascontiguousarray(q[order])
    c = np.ascontiguousarray(costs, dtype=np.float64) if costs is not None else None
    up = np.ascontiguousarray(universe_p, dtype=np.float64) if universe_p is not None else None
    rk = np.ascontiguousarray(ranks, dtype=np.int32) if ranks is not None else None
    out = np.zeros(max(n_sets, 1), dtype=np.int64)
    n = ctypes.c_int64()
    r = lib().orc_set_cover(q.ctypes.data if len(q) else None, len(q), n_sets, n_universes,
                            c.ctypes.data if c is not None else None,
                            up.ctypes.data if up is not None else None,
                            rk.ctypes.data if rk is not None else None,
                            out.ctypes.data, ctypes.byref(n))
    if r != 0:
        raise RuntimeError("oracle set cover failed: %d" % r)
    return [int(x) for x in out[:n.value]]


def approx_multiuniverse(sets, costs=None, universe_p=None, ranks=None):
    """Same call shape as utils/set_cover.py:147 with use_intervalsets=True: `sets` maps
    set_id -> {universe_id: (start,end) | [(start,end), ...] | object with .intervals}.
    Returns a Python set built by .add() in pick order (as the reference builds it)."""
    set_ids = sorted(sets.keys())
    sid = {s: i for i, s in enumerate(set_ids)}
    uids = []
    useen = {}
    for s in set_ids:
        for u in sets[s].keys():
            if u not in useen:
                useen[u] = len(uids)
                uids.append(u)
    quads = []
    for s in set_ids:
        for u, iv in sets[s].items():
            if isinstance(iv, tuple) and len(iv) == 2 and not isinstance(iv[0], tuple):
                ivs = [iv]
            elif hasattr(iv, 'intervals'):
                ivs = list(iv.intervals)
            else:
                ivs = list(iv)
            # merge as interval.IntervalSet does (utils/interval.py:288-316)
            ivs = sorted(ivs)
            merged = []
            for a, b in ivs:
                if merged and a <= merged[-1][1]:
                    merged[-1][1] = max(merged[-1][1], b)
                else:
                    merged.append([a, b])
            for a, b in merged:
                quads.append((sid[s], useen[u], a, b))
    c = [costs[s] for s in set_ids] if costs is not None else None
    up = [universe_p[u] for u in uids] if universe_p is not None else None
    rk = [ranks[s] for s in set_ids] if ranks is not None else None
    picks = set_cover_quads(np.array(quads, dtype=np.int64).reshape(-1, 4), len(set_ids), len(uids),
                            c, up, rk)
    out = set()
    for p in picks:
        out.add(set_ids[p])
    return out


# --------------------------------------------------------------------------- a8-a13
def set_cover_filter(probe_strs_grouped, genomes_grouped, mismatches, lcf_thres,
                     island_of_exact_match=0, coverage=1.0, cover_extension=0, kmer_probe_map_k=20,
                     n_threads=1, return_details=False):
    """filter/set_cover_filter.py:902-930 SetCoverFilter._filter without identify / avoided
    genomes (ranks all 0, :670-735).  genomes_grouped: list (groups) of lists (genomes) of lists
    (sequence strings).  Returns, per group, the selected probe LIST indices in the order the
    reference emits them (iteration order of the pickled-and-restored Python set, :893-900,926)."""
    selected = []
    details = []
    for probes, genomes in zip(probe_strs_grouped, genomes_grouped):
        probes = list(probes)
        if len(probes) == 0:                                   # :393-394
            selected.append([])
            details.append(dict(quads=np.zeros((0, 4), np.int64), picks=[]))
            continue
        k, seeds, _ = choose_seeds(probes, mismatches, lcf_thres, min_k=kmer_probe_map_k,
                                   k=kmer_probe_map_k)
        sm = SeedMap(probes, seeds, k)
        quads = make_sets_quads(sm, genomes, mismatches, lcf_thres, island_of_exact_match,
                                cover_extension, n_threads=n_threads)
        if coverage <= 1.0:                                    # :775-792
            up = [coverage] * len(genomes)
        else:
            up = []
            for g in genomes:
                size = sum(len(s) for s in g)
                up.append(float(min(coverage, size)) / size)
        picks = set_cover_quads(quads, len(probes), len(genomes), None, up,
                                [0] * len(probes))
        s = set()
        for p in picks:
            s.add(p)
        s = pickle.loads(pickle.dumps(s))                      # Pool.starmap round trip
        selected.append([i for i in s])
        details.append(dict(quads=quads, picks=picks))
    return (selected, details) if return_details else selected


# --------------------------------------------------------------------------- a14-a17
def num_tables(P1, k, reporting_prob):
    """utils/lsh.py:270-277."""
    if P1 == 1.0:
        return 1
    return int(math.ceil(math.log(1.0 - reporting_prob, 1.0 - math.pow(P1, k))))


def _priority_order(probe_strs):
    """filter/near_duplicate_filter.py:61-66: distinct probes by multiplicity, descending,
    stable in first-occurrence order."""
    occ = {}
    for p in probe_strs:
        occ[p] = occ.get(p, 0) + 1
    return [p for p, _ in sorted(occ.items(), key=lambda kv: kv[1], reverse=True)]


def _near_dup(order, family, n_tab, k, kmer, pa, pb, dist_thres):
    buf, off = _flatten_strs(order)
    a = np.array(pa, dtype=np.uint64)
    b = np.array(pb, dtype=np.uint64)
    keep = np.zeros(max(len(order), 1), dtype=np.uint8)
    r = lib().orc_near_duplicate(buf.ctypes.data, off.ctypes.data, len(order), family, n_tab, k, kmer,
                                 a.ctypes.data, b.ctypes.data, float(dist_thres), keep.ctypes.data)
    if r != 0:
        raise RuntimeError("oracle near-duplicate failed: %d" % r)
    out = set()
    for i, p in enumerate(order):
        if keep[i]:
            out.add(p)
    return list(out)                                           # near_duplicate_filter.py:103


def near_duplicate_minhash(probe_strs, dist_thres, kmer_size=10, k=3, reporting_prob=0.80):
    """NearDuplicateFilterWithMinHash._filter (filter/near_duplicate_filter.py:159-191).
    Draws (a, b) from Python's `random` exactly as lsh.py:91-96 / :284-287 do; valid for a
    reference run under PYTHONHASHSEED=0."""
    order = _priority_order(probe_strs)
    n_tab = num_tables(1.0 - dist_thres, k, reporting_prob)
    p = 2 ** 31 - 1
    pa, pb = [], []
    for _ in range(n_tab * k):
        pa.append(random.randint(1, p))
        pb.append(random.randint(0, p))
    return _near_dup(order, 0, n_tab, k, kmer_size, pa, pb, dist_thres)


def near_duplicate_hamming(probe_strs, dist_thres, probe_length, k=20, reporting_prob=0.80):
    """NearDuplicateFilterWithHammingDistance._filter (filter/near_duplicate_filter.py:111-146)."""
    order = _priority_order(probe_strs)
    for s in order:
        assert len(s) == probe_length                          # lsh.py:30
    n_tab = num_tables(1.0 - float(dist_thres) / float(probe_length), k, reporting_prob)
    pa = [random.randint(0, probe_length - 1) for _ in range(n_tab * k)]
    return _near_dup(order, 1, n_tab, k, 0, pa, [0] * len(pa), dist_thres)


def abs_pyhash(s):
    return lib().orc_abs_pyhash(_b(s), len(s))


def minhash(s, kmer, a, b):
    return lib().orc_minhash(_b(s), len(s), kmer, a, b)


# ---- genome clustering by MinHash sketches -- small cases only (one hashlib call per k-mer) -------
def sketch_params():
    """The two draws of MinHashFamily.make_h (utils/lsh.py:95-96)."""
    p = 2 ** 31 - 1
    a = random.randint(1, p)
    b = random.randint(0, p)
    return a, b


def sketch(s, kmer_size, N, a, b):
    """h(s) of MinHashFamily(kmer_size, N).make_h() with the md5 inner hash (utils/lsh.py:106-148): the N smallest
    values of (a * int(md5(kmer).hexdigest(), 16) + b) mod (2^31 - 1) over the k-mer list, which is run through
    repeatedly until at least N values have been produced; sorted, repeats kept."""
    import hashlib
    import heapq
    p = 2 ** 31 - 1
    assert kmer_size <= len(s)
    num_kmers = len(s) - kmer_size + 1
    one_pass = [(a * int(hashlib.md5(s[i:i + kmer_size].encode('utf-8')).hexdigest(), 16) + b) % p
                for i in range(num_kmers)]
    values = []
    while len(values) < N:                       # :131-139
        values += one_pass
    if N == 1:
        return (min(values),)
    return tuple(sorted(heapq.nsmallest(N, values)))


def sketch_jaccard_dist(hA, hB, N):
    """MinHashFamily.estimate_jaccard_dist (utils/lsh.py:166-214)."""
    ia = ib = inter = union = 0
    while ia < len(hA) and ib < len(hB):
        if union == N:
            break
        if hA[ia] < hB[ib]:
            ia += 1
        elif hA[ia] > hB[ib]:
            ib += 1
        else:
            inter += 1
            ia += 1
            ib += 1
        union += 1
    return 1.0 - float(inter) / union


def jaccard_dist_from_mash_dist(mash_dist, k):
    """utils/cluster.py:49-71."""
    return 1.0 - 1.0 / (2.0 * np.exp(k * mash_dist) - 1)


def condensed_dist_matrix(sigs, N):
    """utils/cluster.py:103-195 (float32 storage, scipy's condensed order)."""
    n = len(sigs)
    out = np.zeros(n * (n - 1) // 2, dtype=np.float32)
    for j in range(n):
        for i in range(j):
            out[int((-1 * i * i) / 2 + i * n - 3 * i / 2 + j - 1)] = sketch_jaccard_dist(sigs[i], sigs[j], N)
    return out


def connected_components(n, dist_fn, threshold, early_stop_threshold=None):
    """utils/cluster.py:239-355 without the process pool: same search, same set operations."""
    if early_stop_threshold is None:
        early_stop_threshold = jaccard_dist_from_mash_dist(0.02, 12)
    indices_to_consider = set(range(n))

    def dfs(i):
        visited = set()
        to_visit = [i]
        seen = {i}
        while len(to_visit) > 0:
            j = to_visit.pop()
            if j in visited:
                continue
            visited.add(j)
            possible = list(indices_to_consider - seen)
            for k in possible:
                dist = dist_fn(j, k)
                if dist <= threshold:
                    if dist <= early_stop_threshold:
                        visited.add(k)
                        seen.add(k)
                    else:
                        to_visit.append(k)
                        seen.add(k)
        return visited

    previously = set()
    ccs = []
    for i in range(n):
        if i in previously:
            continue
        cc = dfs(i)
        previously.update(cc)
        indices_to_consider -= cc
        ccs.append(sorted(list(cc)))
    ccs.sort(key=len, reverse=True)
    return ccs


def cluster_with_minhash_signatures(seqs, k=12, N=100, threshold=0.1, cluster_method='simple'):
    """utils/cluster.py:358-430; `seqs` is an ordered dict header -> sequence."""
    a, b = sketch_params()
    headers = list(seqs.keys())
    sigs = [sketch(seqs[h], k, N, a, b) for h in headers]
    thr = jaccard_dist_from_mash_dist(threshold, k)
    if cluster_method == 'simple':
        clusters = connected_components(len(sigs), lambda i, j: sketch_jaccard_dist(sigs[i], sigs[j], N), thr)
    else:
        from scipy.cluster import hierarchy
        dm = condensed_dist_matrix(sigs, N)
        if len(dm) == 0:
            clusters = [[0]]
        else:
            lab = hierarchy.fcluster(hierarchy.linkage(dm, method='average'), thr, criterion='distance')
            by = {}
            for i, c in enumerate(lab):
                by.setdefault(int(c), []).append(i)
            sizes = {c: len(by[c]) for c in range(int(min(lab)), int(max(lab)) + 1)}
            clusters = [by[c] for c, _ in sorted(sizes.items(), key=lambda t: t[1], reverse=True)]
    return [[headers[i] for i in c] for c in clusters]
