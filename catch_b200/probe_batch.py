"""Candidate probes as ONE buffer: the batch container that keeps the candidates of a
grouping contiguous from tiling to the set cover.

The reference creates one Python `Probe` object per candidate (filter/candidate_probes.py:21-124,
probe.py:42-53; ~45 us each, 40 million of them at V-All scale) and every filter walks those objects
again.  Here a grouping's candidates are the rows of one [n, L] uint8 array: tiling is a strided
view of the genome bytes, the duplicate / near-duplicate / set-cover filters hand the array to the
device as it is, and `Probe` objects are only materialised for what a caller actually looks at --
in a design run, the selected probes.

A ProbeBatch behaves like a read-only sequence of Probe objects (len, iteration, indexing), so code
that does not know about it still works, just without the speed-up.
"""
import re

import numpy as np

from catch_b200 import probe

_N = ord('N')


class ProbeBatch:
    __slots__ = ('data', 'flanking')

    def __init__(self, data, flanking=None):
        """data: C-contiguous uint8 [n, L] (ASCII); flanking: bool [n] (Probe.is_flanking_n_string) or None."""
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        if self.data.ndim != 2:
            raise ValueError("ProbeBatch needs a [n, L] array")
        self.flanking = flanking

    # ---- the sequence protocol
    def __len__(self):
        return self.data.shape[0]

    @property
    def probe_length(self):
        return self.data.shape[1]

    def probe(self, i):
        p = probe.Probe(self.data[i].tobytes().decode('latin-1'))
        if self.flanking is not None and self.flanking[i]:
            p.is_flanking_n_string = True
        return p

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return self.probe(int(i))
        return self.take(np.arange(len(self))[i])

    def __iter__(self):
        for i in range(len(self)):
            yield self.probe(i)

    def take(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        return ProbeBatch(self.data[idx], None if self.flanking is None else self.flanking[idx])

    def strs(self, idx=None):
        """Sequences (str) of all probes, or of the probes `idx`."""
        rows = self.data if idx is None else self.data[np.asarray(idx, dtype=np.int64)]
        n, L = rows.shape
        if n == 0:
            return []
        flat = rows.tobytes().decode('latin-1')
        return [flat[i * L:(i + 1) * L] for i in range(n)]

    def probes(self, idx):
        """Probe objects of the rows `idx`, built from one decode of the gathered rows."""
        idx = np.asarray(idx, dtype=np.int64)
        out = [probe.Probe(s) for s in self.strs(idx)]
        if self.flanking is not None:
            for p, fl in zip(out, self.flanking[idx].tolist()):
                if fl:
                    p.is_flanking_n_string = True
        return out

    def lengths(self):
        return np.full(len(self), self.data.shape[1], dtype=np.int32)

    @staticmethod
    def concat(batches):
        batches = [b for b in batches if len(b)]
        if not batches:
            return None
        L = batches[0].probe_length
        if any(b.probe_length != L for b in batches):
            raise ValueError("batches of different probe lengths")
        flanking = None
        if any(b.flanking is not None for b in batches):
            flanking = np.concatenate([b.flanking if b.flanking is not None else np.zeros(len(b), dtype=bool)
                                       for b in batches])
        return ProbeBatch(np.concatenate([b.data for b in batches]), flanking)

    # ---- tiling
    @staticmethod
    def from_sequence(seq, probe_length, probe_stride, min_n_string_length=2):
        """filter/candidate_probes.py:21-124 for a sequence at least probe_length long, same probes in the same
        order: tiles at 0, stride, ... while they fit (:97-100); one more flush with the end when
        len % stride != 0 (:102-106); then, per run of >= min_n_string_length N, the probes flanking it
        (:112-122); any probe that itself contains such a run is dropped (:84-85)."""
        n = len(seq)
        L = probe_length
        if n < L:
            raise ValueError("sequence shorter than the probe length: use the per-object path")
        arr = np.frombuffer(seq.encode('latin-1'), dtype=np.uint8)
        starts = np.arange(0, n - L + 1, probe_stride, dtype=np.int64)
        if n % probe_stride != 0:
            starts = np.append(starts, n - L)
        flank = np.zeros(len(starts), dtype=bool)
        run = 'N' * min_n_string_length
        if run in seq:
            extra = []
            for m in re.finditer('(N{%d,})' % min_n_string_length, seq):
                if m.start() - L >= 0:
                    extra.append(m.start() - L)
                if m.end() + L <= n:
                    extra.append(m.end())
            starts = np.concatenate([starts, np.array(extra, dtype=np.int64)])
            flank = np.concatenate([flank, np.ones(len(extra), dtype=bool)])
            # drop every probe that contains a run: run_start[i] = N at i .. i+min-1
            is_n = (arr == _N).astype(np.int32)
            win = np.convolve(is_n, np.ones(min_n_string_length, dtype=np.int32), mode='valid') == min_n_string_length
            pre = np.concatenate([[0], np.cumsum(win)])
            lo, hi = starts, starts + L - min_n_string_length + 1        # run starts inside [lo, hi)
            has_run = (pre[np.maximum(hi, lo)] - pre[lo]) > 0
            starts, flank = starts[~has_run], flank[~has_run]
        windows = np.lib.stride_tricks.sliding_window_view(arr, L)
        return ProbeBatch(windows[starts], flank if flank.any() else None)

    @staticmethod
    def from_sequences(seqs, probe_length, probe_stride, min_n_string_length=2, seq_length_to_skip=None):
        """filter/candidate_probes.py:127-182 for sequences that are all at least probe_length long (after the
        `seq_length_to_skip` rule); returns None when some sequence is shorter -- the caller then takes the
        per-object path, which knows the small-sequence rules.  All sequences without a run of N are tiled in ONE
        pass: the tile starts of every sequence are laid out with a few vector operations over the concatenated
        sequences and the probes are gathered by a single indexed copy (40 000 influenza segments: 1.65 s of
        per-sequence numpy calls -> 0.25 s); a sequence with a run of N goes through from_sequence."""
        L = probe_length
        kept = []
        for s in seqs:
            if seq_length_to_skip is not None and len(s) <= seq_length_to_skip:
                continue
            if len(s) < L:
                return None
            kept.append(s)
        if not kept:
            return ProbeBatch(np.zeros((0, L), dtype=np.uint8))
        run = 'N' * min_n_string_length
        special = [i for i, s in enumerate(kept) if run in s]
        if special:
            # keep the order of the sequences: stretches of plain sequences in one pass each, the others one by one
            parts, lo = [], 0
            for i in special:
                if i > lo:
                    parts.append(ProbeBatch._tile_plain(kept[lo:i], L, probe_stride))
                parts.append(ProbeBatch.from_sequence(kept[i], L, probe_stride, min_n_string_length))
                lo = i + 1
            if lo < len(kept):
                parts.append(ProbeBatch._tile_plain(kept[lo:], L, probe_stride))
            out = ProbeBatch.concat(parts)
            return out if out is not None else ProbeBatch(np.zeros((0, L), dtype=np.uint8))
        return ProbeBatch._tile_plain(kept, L, probe_stride)

    @staticmethod
    def _tile_plain(seqs, L, stride):
        """Tiles of sequences that hold no run of N (so no probe is dropped and none is added, :97-106): for a
        sequence of length n the starts 0, stride, ... while start + L <= n, then n - L when n % stride != 0."""
        lens = np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs))
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        arr = np.frombuffer(''.join(seqs).encode('latin-1'), dtype=np.uint8)
        regular = (lens - L) // stride + 1
        counts = regular + (lens % stride != 0)
        total = int(counts.sum())
        seq_id = np.repeat(np.arange(len(seqs), dtype=np.int64), counts)
        first = np.zeros(len(seqs), dtype=np.int64)
        np.cumsum(counts[:-1], out=first[1:])
        within = np.arange(total, dtype=np.int64) - first[seq_id]
        starts = np.where(within < regular[seq_id], within * stride, lens[seq_id] - L) + off[seq_id]
        windows = np.lib.stride_tricks.sliding_window_view(arr, L)
        return ProbeBatch(windows[starts])
