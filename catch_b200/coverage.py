"""Host side of stage A: packs a group's probes and target genomes, replays the seed choice,
and drives cb_coverage / cb_setcover.  Everything numeric happens in libcatchb200.so."""
import itertools

import numpy as np

import ctypes

from catch_b200 import _lib
from catch_b200 import probe as probe_mod
from catch_b200.probe_batch import ProbeBatch

try:                                    # host-side glue built next to libcatchb200.so (csrc/fastpack.c)
    from catch_b200 import _fastpack
except ImportError:                     # pure-Python gathering below does the same, slower
    _fastpack = None


def _concat_ascii(strs):
    """(uint8 array of all bytes, int64 offsets) for a list of str."""
    n = len(strs)
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter((len(s) for s in strs), dtype=np.int64, count=n), out=off[1:])
    raw = ''.join(strs).encode('latin-1')
    if len(raw) != off[-1]:
        raise ValueError("sequences must contain single-byte characters only")
    buf = np.frombuffer(raw if raw else b'\0', dtype=np.uint8)
    return buf, off, raw


def gather_staged(ctx, slot, items):
    """Gather the sequences of `items` (str or objects with .seq_str) straight into the context's
    page-locked staging buffer `slot`.  Returns (address, int32 lengths, total bytes), or None when
    the C helper is not built (callers then use gather_probes)."""
    if isinstance(items, ProbeBatch):             # already one buffer: a single copy into the staging memory
        if not hasattr(ctx, 'host_buffer'):
            return None
        total = int(items.data.size)
        addr, cap = ctx.host_buffer(slot, total)
        if total:
            data = items.data if items.data.flags['C_CONTIGUOUS'] else np.ascontiguousarray(items.data)
            if _fastpack is not None and hasattr(_fastpack, 'copy_into'):
                _fastpack.copy_into(data, addr)          # threaded from 24 MB on, GIL released
            else:
                ctypes.memmove(addr, data.ctypes.data, total)
        return addr, items.lengths(), total
    if _fastpack is None or not hasattr(ctx, 'host_buffer'):
        return None
    n = len(items)
    addr, cap = ctx.host_buffer(slot, 0)
    total, lens = _fastpack.gather_into(items, 'seq_str', addr, cap)
    if total > cap:
        addr, cap = ctx.host_buffer(slot, total)
        total, lens = _fastpack.gather_into(items, 'seq_str', addr, cap)
    return addr, np.frombuffer(lens, dtype=np.int32, count=n), total


def fingerprint(gathered, max_samples=64):
    """Cheap checksum of a gathered probe list (number of probes, their lengths' sum and the bytes of up to
    `max_samples` evenly spaced probes): ranks that are about to shard ONE probe list compare it, because a list whose
    order came out of a Python set differs between processes unless PYTHONHASHSEED is pinned."""
    import zlib
    raw, lens = gathered[0], gathered[1]
    n = len(lens)
    if n == 0:
        return 0
    off = offsets_from_lengths(lens)
    crc = zlib.crc32(b'%d:%d' % (n, int(off[-1])))
    for i in np.unique(np.linspace(0, n - 1, num=min(n, max_samples)).astype(np.int64)).tolist():
        a, b = int(off[i]), int(off[i + 1])
        chunk = raw[a:b] if isinstance(raw, (bytes, bytearray)) else ctypes.string_at(raw + a, b - a)
        crc = zlib.crc32(chunk, crc)
    return crc


def fingerprint_lists(groupings, per_list=16):
    """The same idea for a list of probe lists that are NOT gathered on this rank (group-sharded runs gather only
    the groupings a rank owns): lengths of the lists and up to `per_list` evenly spaced sequences of each."""
    import zlib
    crc = 0
    for probes in groupings:
        if not isinstance(probes, (list, tuple, ProbeBatch)):
            continue                                     # some other iterable: nothing to sample without consuming it
        n = len(probes)
        crc = zlib.crc32(b'%d;' % n, crc)
        if n == 0:
            continue
        idx = sorted(set(int(round(x)) for x in np.linspace(0, n - 1, num=min(n, per_list)).tolist()))
        if isinstance(probes, ProbeBatch):
            crc = zlib.crc32(probes.data[np.asarray(idx, dtype=np.int64)].tobytes(), crc)
        else:
            for i in idx:
                p = probes[i]
                crc = zlib.crc32((p if isinstance(p, str) else p.seq_str).encode('latin-1'), crc)
    return crc


def offsets_from_lengths(lens):
    """int64 offsets [n + 1] of sequences of the given lengths laid back to back."""
    n = len(lens)
    if n and int(lens[0]) > 0 and bool((lens == lens[0]).all()):       # candidate probes: one length
        return np.arange(n + 1, dtype=np.int64) * int(lens[0])
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, dtype=np.int64, out=off[1:])
    return off


def probe_lengths(probes):
    """int32 lengths of the sequences of `probes` (str or objects with .seq_str), nothing copied."""
    if isinstance(probes, ProbeBatch):
        return probes.lengths()
    n = len(probes)
    if _fastpack is not None:
        _, lens = _fastpack.lengths(probes, 'seq_str')
        return np.frombuffer(lens, dtype=np.int32, count=n)
    return np.fromiter((len(p if isinstance(p, str) else p.seq_str) for p in probes), dtype=np.int32, count=n)


def gather_probes(probes):
    """(bytes of all sequences back to back, int32 lengths) for a list of Probe objects (or str).
    One C pass over the list when the _fastpack helper is built."""
    if isinstance(probes, ProbeBatch):
        return probes.data.tobytes(), probes.lengths()
    n = len(probes)
    if _fastpack is not None:
        data, lens = _fastpack.gather(probes, 'seq_str')
        return data, np.frombuffer(lens, dtype=np.int32, count=n)
    strs = [p if isinstance(p, str) else p.seq_str for p in probes]
    lens = np.fromiter(map(len, strs), dtype=np.int32, count=n)
    try:
        data = ''.join(strs).encode('latin-1')
    except UnicodeEncodeError:
        raise ValueError("sequences must contain single-byte characters only")
    return data, lens


def stage_targets(ctx, slot, genomes):
    """Gather the sequences of a grouping's genomes into staging buffer `slot` (or a bytes object when the C
    helper is not built) and derive the tables cb_upload_group needs.  No library call is made when the
    buffer is already large enough, so a helper thread may run this while the context is busy."""
    # sequences of all genomes in order, and the genome each belongs to (no per-sequence Python loop:
    # an influenza-shaped grouping has 40 000 single-sequence genomes)
    seq_lists = [g.seqs if hasattr(g, 'seqs') else g for g in genomes]
    n_genomes = len(genomes)
    counts = np.fromiter(map(len, seq_lists), dtype=np.int64, count=n_genomes)
    seqs = list(itertools.chain.from_iterable(seq_lists))
    sg = np.repeat(np.arange(n_genomes, dtype=np.int32), counts) if len(seqs) else np.zeros(1, np.int32)
    staged_t = gather_staged(ctx, slot, seqs)
    if staged_t is not None:
        t_raw, t_lens, t_total = staged_t
        seq_off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(t_lens, dtype=np.int64, out=seq_off[1:])
    else:
        _, seq_off, t_raw = _concat_ascii(seqs)
        t_total = len(t_raw)
    return dict(raw=t_raw, seq_off=seq_off, seq_genome=sg, n_genomes=n_genomes, total=int(t_total))


class PackedGroup:
    """Probes + targets of one grouping, resident on the device (one cb_upload_group call: both
    host->device copies, the code table derived on the device, both packings)."""

    def __init__(self, ctx, probe_strs, genomes, gathered=None, targets_staged=None):
        """`probe_strs`: list of str (or of objects with .seq_str).  `gathered`: the result of
        gather_probes() / gather_staged() on it, `targets_staged`: the result of stage_targets() on
        `genomes`, if the caller already has them (prefetched while the previous grouping was on the device)."""
        self.ctx = ctx
        self.n_probes = len(probe_strs)
        ts = targets_staged if targets_staged is not None else stage_targets(ctx, 1, genomes)
        self.n_genomes = ts['n_genomes']
        self.target_bases = int(ts['seq_off'][-1])
        if gathered is None:
            gathered = gather_staged(ctx, 0, probe_strs)
            if gathered is None:
                gathered = gather_probes(probe_strs)
        p_raw, lens = gathered[0], gathered[1]
        off = offsets_from_lengths(lens)
        self.probes, self.targets, self.probe_len, self.bits, st = ctx.upload_group(
            p_raw, self.n_probes, ts['raw'], ts['seq_off'], ts['seq_genome'], self.n_genomes, probe_off=off)
        self.st_targets, self.st_probes = st, _lib.Stats()
        self.h2d_bytes = int(off[-1]) + ts['total']

    @property
    def probe_off(self):
        off = np.zeros(self.n_probes + 1, dtype=np.int64)
        np.cumsum(self.probe_len, out=off[1:])
        return off

    def free(self):
        self.targets.free()
        self.probes.free()


def seeds_to_csr(seeds, rep=None):
    """[n, s] seed draws -> CSR for cb_coverage (set semantics on the device, so repeats are
    passed through untouched).  `rep[i]` (optional) redirects the draws of list index i to another
    index: duplicates collapse onto their last occurrence, see dedup_map."""
    n = seeds.shape[0]
    if n == 0:
        return np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32)
    if rep is None:
        s = seeds.shape[1]
        off = np.arange(n + 1, dtype=np.int64) * s
        pos = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1)
    else:
        merged = [[] for _ in range(n)]
        for i in range(n):
            merged[rep[i]].extend(int(x) for x in seeds[i])
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(m) for m in merged])
        pos = np.array([x for m in merged for x in m], dtype=np.int32)
    if pos.size == 0:
        pos = np.zeros(1, dtype=np.int32)
    return off, pos


def dedup_map(probe_strs):
    """For every list index the LAST index holding the same sequence
    (filter/set_cover_filter.py:408-412: probe_id[p] = id overwrites earlier duplicates, and the
    k-mer map is keyed by sequence, probe.py:324-329).  Returns None when all are distinct."""
    if len(set(probe_strs)) == len(probe_strs):
        return None
    last = {}
    for i, s in enumerate(probe_strs):
        last[s] = i
    return [last[s] for s in probe_strs]


def draw_seeds(lengths, mismatches, lcf_thres, kmer_probe_map_k, background=False):
    """(k, seeds, mode) for probes of the given lengths; consumes numpy's global RNG in random
    mode.  Only needs the lengths.  With background=True the generator may run on the library's
    worker thread while the caller packs and uploads the sequences; `seeds` is then a pending
    object until finish_draw() is applied (nothing else may touch np.random in between)."""
    return probe_mod.choose_seed_positions(lengths, mismatches, lcf_thres, min_k=kmer_probe_map_k,
                                           k=kmer_probe_map_k, randint=_lib.legacy_randint,
                                           randint_async=_lib.PendingRandint if background else None)


def cancel_draw(drawn):
    """Drop a background draw (made on a guess that turned out wrong) without consuming the RNG."""
    seeds = drawn[1]
    if hasattr(seeds, 'cancel'):
        seeds.cancel()


def finish_draw(drawn):
    k, seeds, mode = drawn
    if hasattr(seeds, 'result'):
        seeds = seeds.result()
    return k, seeds, mode


class SeedPlan:
    """The seed choice for one probe list: drawn on the host (it consumes numpy's global RNG
    exactly like probe.construct_kmer_probe_map_to_find_probe_covers, probe.py:507-577), then
    handed to cb_coverage as a CSR.  Drawing is separate from the device call so that a rank
    which does not own a grouping can still advance the RNG stream identically."""

    def __init__(self, probe_strs, mismatches, lcf_thres, kmer_probe_map_k, lengths=None, may_have_dups=True,
                 drawn=None):
        """`drawn`: the result of draw_seeds() when the draw was done ahead of time."""
        if drawn is None:
            if lengths is None:
                lengths = np.fromiter(map(len, probe_strs), dtype=np.int64, count=len(probe_strs))
            drawn = draw_seeds(lengths, mismatches, lcf_thres, kmer_probe_map_k)
        self.k, seeds, self.mode = drawn
        # may_have_dups=False: the device already established that all probes are distinct
        self.rep = dedup_map(probe_strs) if may_have_dups else None
        seeds = np.asarray(seeds)
        # Without duplicates every probe has the same number of draws: the dense byte matrix goes to
        # cb_coverage_uniform as it is.  With duplicates the draws of a sequence's copies are merged
        # (the k-mer map is keyed by sequence) and the general CSR form is needed.
        self.uniform = None
        self._seeds = seeds
        self._csr = None
        if self.rep is None and seeds.ndim == 2 and seeds.size and \
                (seeds.dtype == np.uint8 or int(seeds.max(initial=0)) < 256):
            self.uniform = np.ascontiguousarray(seeds, dtype=np.uint8)

    def _get_csr(self):
        if self._csr is None:
            self._csr = seeds_to_csr(self._seeds, self.rep)
        return self._csr

    @property
    def seed_off(self):
        return self._get_csr()[0]

    @property
    def seed_pos(self):
        return self._get_csr()[1]


def compute_cover(ctx, group, plan, mismatches, lcf_thres, island, cover_extension):
    """Stage A for one packed group with a drawn SeedPlan.  Returns (cover handle, stats)."""
    if plan.uniform is not None:
        return ctx.coverage_uniform(group.probes, group.targets, mismatches, lcf_thres, island, cover_extension,
                                    plan.k, plan.uniform)
    return ctx.coverage(group.probes, group.targets, mismatches, lcf_thres, island, cover_extension,
                        plan.k, plan.seed_off, plan.seed_pos)


def compute_cover_range(ctx, group, plan, mismatches, lcf_thres, island, cover_extension, lo, hi):
    """Stage A for the probes [lo, hi) of a packed group (a rank's shard of a probe-sharded run); the
    cover keeps the global probe ids.  `plan` is the seed plan of the WHOLE probe list."""
    if plan.uniform is not None:
        return ctx.coverage_range(group.probes, group.targets, mismatches, lcf_thres, island, cover_extension,
                                  plan.k, lo, hi, seeds_u8=plan.uniform[lo:hi])
    return ctx.coverage_range(group.probes, group.targets, mismatches, lcf_thres, island, cover_extension,
                              plan.k, lo, hi, seed_off=plan.seed_off, seed_pos=plan.seed_pos)


class KmerMapOrder:
    """Order in which the reference's k-mer map lists the probes that share one k-mer.  The map's values
    are Python SETS of (probe, position) tuples (probe.py:385, :497), filled probe by probe in list order
    and k-mer by k-mer in draw order, and find_probe_covers_in_sequence aligns the probes of a k-mer in the
    iteration order of that set (probe.py:756-760, :1067).  That order decides which of several probes first
    seen at the SAME sequence position comes first in the scan's result dict, which the consumers with
    order-dependent output inherit (interval.schedule ties in the adapter filter, the row order of the
    analyzer's probe-map-counts file).  It is reproduced with the real thing: the same insertions into real
    sets of (Probe, position) tuples -- Probe hashes by its sequence, so like the reference this is
    reproducible across processes only under a fixed PYTHONHASHSEED.  Built lazily: ties are rare."""

    def __init__(self, probes, plan):
        self.probes, self.plan, self.sets = probes, plan, None

    def _build(self):
        k, seeds = self.plan.k, np.asarray(self.plan._seeds)
        sets = {}
        for i, p in enumerate(self.probes):
            s = p.seq_str
            for pos in seeds[i].tolist():
                sets.setdefault(s[pos:pos + k], set()).add((p, pos))
        self.sets = sets

    def rank(self, kmer, probe):
        """Index of `probe`'s first entry in the iteration order of the k-mer's set (large if absent)."""
        if self.sets is None:
            self._build()
        for r, (q, _pos) in enumerate(self.sets.get(kmer, ())):
            if q == probe:
                return r
        return 1 << 30


def listing_tie_ranks(g_probe, g_hit, sequence, kmer_order):
    """Secondary sort key of the listing order: 0 everywhere except for probes whose first hit is at the same
    sequence position as another probe's; those get their rank in the k-mer map's set of that position's k-mer."""
    tie = np.zeros(len(g_probe), dtype=np.int64)
    if kmer_order is None or sequence is None or len(g_probe) < 2:
        return tie
    probes_u, first = np.unique(g_probe, return_index=True)
    hits_u = g_hit[first]
    vals, counts = np.unique(hits_u, return_counts=True)
    for h in vals[counts > 1].tolist():
        kmer = sequence[h:h + kmer_order.plan.k]
        for pi in probes_u[hits_u == h].tolist():
            tie[g_probe == pi] = kmer_order.rank(kmer, kmer_order.probes[pi])
    return tie


def sequence_batches(sequences, max_bases=1 << 30):
    """Split an iterable of sequences into lists whose total length stays below max_bases, so that a
    grouping of any size goes through the device in bounded pieces (a universe is limited to 2^32 bits,
    cb_upload_targets).  A single sequence longer than max_bases forms its own batch."""
    batch, total = [], 0
    for s in sequences:
        if batch and total + len(s) > max_bases:
            yield batch
            batch, total = [], 0
        batch.append(s)
        total += len(s)
    if batch:
        yield batch


def scan_records(ctx, probe_strs, sequences, plan, mismatches, lcf_thres, island, max_bases=1 << 30):
    """The device counterpart of probe.find_probe_covers_in_sequence() over every sequence of `sequences`
    BEFORE any merging (probe.py:1008-1119): an int64 array [n, 5] of (probe index, sequence index, start,
    end, position of the seed hit) with one row per emitted range, in no particular order.  The two uses
    in the reference: merge_overlapping=False callers take sorted(set(ranges)) per probe and sequence
    (probe.py:1262-1270), merge_overlapping=True callers merge them (utils/interval.py:288-316); the hit
    position gives the order in which the reference's result dict first sees each probe."""
    out, base = [], 0
    seqs = list(sequences)
    for batch in sequence_batches(seqs, max_bases):
        group = PackedGroup(ctx, probe_strs, [[s] for s in batch])
        try:
            rec, _ = ctx.coverage_records(group.probes, group.targets, mismatches, lcf_thres, island, plan.k,
                                          plan.seed_off, plan.seed_pos)
        finally:
            group.free()
        if len(rec):
            rec[:, 1] += base
            out.append(rec)
        base += len(batch)
    return np.concatenate(out) if out else np.zeros((0, 5), dtype=np.int64)


def cover_with_seeds(ctx, group, seeds_per_probe, k, mismatches, lcf_thres, island, cover_extension):
    """Stage A with explicitly given seed positions (a list of position lists, one per probe):
    the device counterpart of probe.find_probe_covers_in_sequence over a prebuilt
    kmer_probe_map (probe.py:1122)."""
    n = len(seeds_per_probe)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in seeds_per_probe])
    flat = [int(x) for sl in seeds_per_probe for x in sl]
    pos = np.array(flat if flat else [0], dtype=np.int32)
    return ctx.coverage(group.probes, group.targets, mismatches, lcf_thres, island, cover_extension, k,
                        off, pos)
