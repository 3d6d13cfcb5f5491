"""Host-side probe type and seed selection.

Mirrors the parts of the reference's catch/probe.py that the hot path's callers touch:
`Probe` (probe.py:38-354; equality and hashing by sequence, :324-329) and the choice of seed
k-mers of construct_kmer_probe_map_to_find_probe_covers (probe.py:507-577).  The scan itself
(probe.py:1008-1271) runs on the device; see csrc/coverage.cu.
"""
import hashlib

import numpy as np

_RC = {'A': 'T', 'T': 'A', 'C': 'G', 'G': 'C'}


class Probe:
    """Immutable probe sequence.  `seq` (numpy array of single characters, as in the reference)
    is materialised lazily; the device path only ever needs `seq_str`."""

    __slots__ = ('seq_str', '_seq', 'is_flanking_n_string', 'header')

    def __init__(self, seq):
        if isinstance(seq, str):
            self.seq_str = seq
            self._seq = None
        else:
            self._seq = seq
            self.seq_str = ''.join(seq)
        self.is_flanking_n_string = False
        self.header = None

    @property
    def seq(self):
        if self._seq is None:
            self._seq = np.fromiter(self.seq_str, dtype='U1', count=len(self.seq_str))
        return self._seq

    @staticmethod
    def from_str(seq_str):
        return Probe(seq_str)

    def mismatches(self, other):
        return self.mismatches_at_offset(other, 0)

    def mismatches_at_offset(self, other, offset):
        a, b = self.seq_str, other.seq_str
        if len(a) != len(b):
            raise ValueError("Sequences must be of same length")
        if abs(offset) >= len(b):
            raise ValueError("Invalid offset value " + str(offset))
        if offset < 0:
            a, b = a[:offset], b[-offset:]
        elif offset > 0:
            a, b = a[offset:], b[:-offset]
        return sum(1 for x, y in zip(a, b) if x != y)

    def reverse_complement(self):
        return Probe(''.join(_RC.get(c, c) for c in reversed(self.seq_str)))

    def with_prepended_str(self, s):
        return Probe(s + self.seq_str)

    def with_appended_str(self, s):
        return Probe(self.seq_str + s)

    def construct_kmers(self, k, include_positions=False):
        s = self.seq_str
        if include_positions:
            return [(s[i:i + k], i) for i in range(len(s) - k + 1)]
        return [s[i:i + k] for i in range(len(s) - k + 1)]

    def identifier(self, length=10):
        """Last `length` hex digits of the SHA-224 of the sequence (probe.py:301-322); this is
        the FASTA header the reference writes, so it must stay byte-compatible."""
        return hashlib.sha224(self.seq_str.encode()).hexdigest()[-length:]

    def __hash__(self):
        return hash(self.seq_str)

    def __eq__(self, other):
        return hasattr(other, 'seq_str') and self.seq_str == other.seq_str

    def __len__(self):
        return len(self.seq_str)

    def __getitem__(self, i):
        return self.seq_str[i]

    def __str__(self):
        return self.seq_str

    __repr__ = __str__


class PigeonholeRequiresTooSmallKmerSizeError(Exception):
    """probe.py:408-411."""


def pigeonhole_kmer_length(probe_length, mismatches, min_k):
    """Largest k dividing probe_length with probe_length/k > mismatches (probe.py:473-491)."""
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


def choose_seed_positions(lengths, mismatches, lcf_thres, min_k=20, k=20, num_kmers_per_probe=20,
                          randint=None, randint_async=None):
    """Seed start positions for every probe of a list, as the reference would select them.

    Args:
        lengths: int array, length of every probe in LIST ORDER
    Returns:
        (k, seeds, mode): `seeds` is an int array [n_probes, s] of start positions (rows may
        hold repeats; callers apply set semantics), mode is 'pigeonhole' or 'random'.

    Pigeonhole mode is used iff all probes have one length L, mismatches/lcf_thres are given
    and lcf_thres >= L (probe.py:562-573); it falls back to random mode when it would need
    k < min_k (:574-577).  Random mode draws, per probe in list order,
    np.random.choice(L - k + 1, size=20, replace=True) from numpy's legacy global stream
    (probe.py:386-398).  For a run of probes of equal length that is the same stream as one
    np.random.randint(0, L - k + 1, size=(run, 20)) call, which is what is used here; `randint`
    may supply a faster generator of the same stream (catch_b200._lib.legacy_randint);
    `randint_async(bound, shape)` may start that generator in the background, in which case
    `seeds` is returned as a pending object with a result() method (single-length lists only).
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    if n == 0:
        return k, np.zeros((0, 0), dtype=np.int32), 'empty'
    L0 = int(lengths[0])
    differ = bool(np.any(lengths != L0))
    if not (mismatches is None or lcf_thres is None or differ or lcf_thres < L0):
        try:
            kk = pigeonhole_kmer_length(L0, mismatches, min_k)
            row = np.arange(0, L0, kk, dtype=np.int32)
            return kk, np.broadcast_to(row, (n, len(row))), 'pigeonhole'
        except PigeonholeRequiresTooSmallKmerSizeError:
            pass
    if np.any(lengths < k):
        raise ValueError("k is larger than the length of a probe")
    if randint_async is not None and not differ:
        # one run of equal lengths = one generator call, which may run in the background:
        # the third element is then an object whose result() yields the [n, s] draws
        return k, randint_async(L0 - k + 1, (n, num_kmers_per_probe)), 'random'
    seeds = np.empty((n, num_kmers_per_probe), dtype=np.int32 if int(lengths.max()) - k + 1 > 256 else np.uint8)
    # runs of equal length, in list order
    change = np.flatnonzero(np.diff(lengths)) + 1
    starts = np.concatenate(([0], change))
    ends = np.concatenate((change, [n]))
    for s, e in zip(starts, ends):
        span = int(lengths[s]) - k + 1
        if randint is None:
            seeds[s:e] = np.random.randint(0, span, size=(e - s, num_kmers_per_probe))
        else:
            seeds[s:e] = randint(span, (e - s, num_kmers_per_probe))
    return k, seeds, 'random'
