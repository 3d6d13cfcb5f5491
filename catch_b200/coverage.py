"""Host side of stage A: packs a group's probes and target genomes, replays the seed choice,
and drives cb_coverage / cb_setcover.  Everything numeric happens in libcatchb200.so."""
import numpy as np

from catch_b200 import _lib
from catch_b200 import probe as probe_mod


def _concat_ascii(strs):
    """(uint8 array of all bytes, int64 offsets) for a list of str."""
    n = len(strs)
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter((len(s) for s in strs), dtype=np.int64, count=n), out=off[1:])
    buf = np.frombuffer(''.join(strs).encode('latin-1'), dtype=np.uint8)
    if buf.size != off[-1]:
        raise ValueError("sequences must contain single-byte characters only")
    if buf.size == 0:
        buf = np.zeros(1, dtype=np.uint8)
    return buf, off


def make_alphabet(*byte_arrays):
    """Code table for the bit-plane packing: distinct bytes -> dense codes.  Two bases match
    iff their bytes are equal (utils/longest_common_substring.py:110), so any injective code
    works; ACGT-only input gets 2 planes, ACGT+N 3, arbitrary test alphabets up to 8."""
    present = np.zeros(256, dtype=bool)
    for a in byte_arrays:
        if a.size:
            present |= np.bincount(a, minlength=256).astype(bool)
    for c in b'ACGT':
        present[c] = True
    symbols = np.flatnonzero(present)
    lut = np.zeros(256, dtype=np.uint8)
    lut[symbols] = np.arange(len(symbols), dtype=np.uint8)
    bits = max(1, int(np.ceil(np.log2(len(symbols)))))
    return lut, bits


class PackedGroup:
    """Probes + targets of one grouping, resident on the device."""

    def __init__(self, ctx, probe_strs, genomes):
        self.ctx = ctx
        self.n_probes = len(probe_strs)
        seqs, seq_genome = [], []
        for j, g in enumerate(genomes):
            for s in (g.seqs if hasattr(g, 'seqs') else g):
                seqs.append(s)
                seq_genome.append(j)
        self.n_genomes = len(genomes)
        self.target_bases = sum(len(s) for s in seqs)
        p_buf, self.probe_off = _concat_ascii(probe_strs)
        t_buf, seq_off = _concat_ascii(seqs)
        lut, bits = make_alphabet(p_buf[:self.probe_off[-1]], t_buf[:seq_off[-1]])
        self.bits = bits
        sg = np.array(seq_genome if seq_genome else [0], dtype=np.int32)
        self.targets, self.st_targets = ctx.upload_targets(t_buf, seq_off, sg, self.n_genomes, lut, bits)
        self.probes, self.st_probes = ctx.upload_probes(p_buf, self.probe_off, lut, bits)
        self.h2d_bytes = int(self.probe_off[-1] + seq_off[-1])

    def free(self):
        self.targets.free()
        self.probes.free()


def seeds_to_csr(seeds, rep=None):
    """[n, s] seed draws -> CSR of distinct ascending positions per probe.  `rep[i]` (optional)
    redirects the draws of list index i to another index (duplicates collapse onto the last
    occurrence, see SetCoverFilter._dedup_map)."""
    n = seeds.shape[0]
    if n == 0:
        return np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32)
    if rep is None:
        srt = np.sort(seeds, axis=1)
        keep = np.ones(srt.shape, dtype=bool)
        keep[:, 1:] = srt[:, 1:] != srt[:, :-1]
        counts = keep.sum(axis=1)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        pos = np.ascontiguousarray(srt[keep], dtype=np.int32)
    else:
        merged = [set() for _ in range(n)]
        for i in range(n):
            merged[rep[i]].update(int(x) for x in seeds[i])
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(m) for m in merged])
        pos = np.array([x for m in merged for x in sorted(m)], dtype=np.int32)
    if pos.size == 0:
        pos = np.zeros(1, dtype=np.int32)
    return off, pos


def dedup_map(probe_strs):
    """For every list index the LAST index holding the same sequence
    (filter/set_cover_filter.py:408-412: probe_id[p] = id overwrites earlier duplicates, and the
    k-mer map is keyed by sequence, probe.py:324-329).  Returns None when all are distinct."""
    last = {}
    for i, s in enumerate(probe_strs):
        last[s] = i
    if len(last) == len(probe_strs):
        return None
    return [last[s] for s in probe_strs]


def compute_cover(ctx, group, probe_strs, mismatches, lcf_thres, island, cover_extension,
                  kmer_probe_map_k):
    """Stage A for one packed group.  Returns (cover handle, stats, k, mode)."""
    lengths = np.diff(group.probe_off)
    k, seeds, mode = probe_mod.choose_seed_positions(lengths, mismatches, lcf_thres,
                                                     min_k=kmer_probe_map_k, k=kmer_probe_map_k)
    rep = dedup_map(probe_strs)
    seed_off, seed_pos = seeds_to_csr(np.asarray(seeds), rep)
    cover, st = ctx.coverage(group.probes, group.targets, mismatches, lcf_thres, island,
                             cover_extension, k, seed_off, seed_pos)
    return cover, st, k, mode
