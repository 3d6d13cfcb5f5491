"""Coverage analysis of a probe set on the device scan: the drop-in for catch/coverage_analysis.py.

Same `Analyzer` constructor, `run()`, result attributes (`target_covers`, `bp_covered`,
`average_coverage`, `sliding_coverage`, `probe_map_counts`) and writers as the reference
(coverage_analysis.py:70-600).  What the reference spends its time on -- one
probe.find_probe_covers_in_sequence(sequence, merge_overlapping=False) per target sequence and per
reverse complement (:183-253) -- is one cb_coverage_records call per batch of sequences; the
aggregation per genome is cheap and stays on the host.

merge_overlapping=False semantics (probe.py:1262-1270): per probe and sequence the DISTINCT ranges,
sorted; each is then extended by cover_extension and clipped to the sequence (:236-243).  Order-
dependent outputs follow the reference's dict order: probes in the order the scan first finds them
(see catch_b200/coverage.py: KmerMapOrder).
"""
import logging
from collections import OrderedDict, defaultdict

import numpy as np

from catch_b200 import _lib
from catch_b200 import coverage as cov

logger = logging.getLogger(__name__)

_RC = str.maketrans('ACGT', 'TGCA')


def _table(data, col_justify, header_underline=True):
    """utils/pretty_print.py:7-84 (plain text table, multi-line cells)."""
    if len(data) == 0:
        return ''
    num_cols = len(data[0])
    for row in data:
        if len(row) != num_cols:
            raise ValueError("data has inconsistent number of columns")
    if len(col_justify) != num_cols:
        raise ValueError("col_justify has incorrect number of entries")
    cells = [[str(c).rstrip().split('\n') for c in row] for row in data]
    widths = [max(max(len(line) for line in row[j]) for row in cells) for j in range(num_cols)]
    pad = {'left': str.ljust, 'right': str.rjust, 'center': str.center}
    out = ''
    for i, row in enumerate(cells):
        for h in range(max(len(c) for c in row)):
            parts = []
            for j, c in enumerate(row):
                if col_justify[j] not in pad:
                    raise ValueError("Unknown column justification at %d" % j)
                parts.append(pad[col_justify[j]](c[h] if h < len(c) else '', widths[j]))
            out += ' '.join(parts) + '\n'
        if i == 0 and header_underline:
            out += ' '.join('-' * w for w in widths) + '\n'
    return out


class Analyzer:
    def __init__(self, probes, mismatches, lcf_thres, target_genomes, target_genomes_names=None,
                 island_of_exact_match=0, custom_cover_range_fn=None, cover_extension=0,
                 kmer_probe_map_k=10, rc_too=True):
        if custom_cover_range_fn is not None:
            raise NotImplementedError("custom hybridization functions are Python callables and "
                                      "are not supported by the device implementation")
        self.probes = probes
        self.target_genomes = target_genomes
        if target_genomes_names:
            if len(target_genomes_names) != len(target_genomes):
                raise ValueError(("Number of target genome names must be same "
                                  "as the number of target genomes"))
            self.target_genomes_names = target_genomes_names
        else:
            self.target_genomes_names = ["Group %d" % i for i in range(len(target_genomes))]
        self.mismatches = mismatches
        self.lcf_thres = lcf_thres
        self.island_of_exact_match = island_of_exact_match
        self.cover_extension = cover_extension
        self.kmer_probe_map_k = kmer_probe_map_k
        self.rc_too = rc_too
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _iter_target_genomes(self):
        """coverage_analysis.py:163-181."""
        for i, genomes_from_group in enumerate(self.target_genomes):
            for j, gnm in enumerate(genomes_from_group):
                yield i, j, gnm, False
                if self.rc_too:
                    yield i, j, gnm, True

    # ------------------------------------------------------------------ the scan
    def _find_covers_in_target_genomes(self):
        """coverage_analysis.py:183-253: self.target_covers[i][j][rc] = list of (start, end) in genome
        coordinates (one per probe hybridisation, duplicates kept), self.probe_map_counts[probe] = number
        of sequences (not counting reverse complements) the probe maps to."""
        ctx = self._context()
        probes = list(self.probes)
        probe_strs = [p.seq_str for p in probes]
        self.target_covers = {}
        self.probe_map_counts = OrderedDict()         # a Counter in the reference; insertion order matters
        jobs = []                                      # (i, j, rc, sequence, offset in the genome)
        for i, j, gnm, rc in self._iter_target_genomes():
            self.target_covers.setdefault(i, {}).setdefault(j, {False: None, True: None})
            self.target_covers[i][j][rc] = []
            length_so_far = 0
            for sequence in gnm.seqs:
                if rc:
                    sequence = sequence[::-1].translate(_RC)      # rc_map.get(b, b), :219-222
                jobs.append((i, j, rc, sequence, length_so_far))
                length_so_far += len(sequence)
        if not probes or not jobs:
            return
        plan = cov.SeedPlan(probe_strs, self.mismatches, self.lcf_thres, self.kmer_probe_map_k)
        kmer_order = cov.KmerMapOrder(probes, plan)
        ext = self.cover_extension
        start = 0
        for batch in cov.sequence_batches((job[3] for job in jobs), 1 << 28):
            rec = cov.scan_records(ctx, probe_strs, batch, plan, self.mismatches, self.lcf_thres,
                                   self.island_of_exact_match)
            # distinct (probe, sequence, start, end); first hit of each probe in each sequence
            order = np.lexsort((rec[:, 3], rec[:, 2], rec[:, 0], rec[:, 1]))
            rec = rec[order]
            bounds = np.searchsorted(rec[:, 1], np.arange(len(batch) + 1))
            for q in range(len(batch)):
                i, j, rc, sequence, offset = jobs[start + q]
                r = rec[bounds[q]:bounds[q + 1]]
                if len(r) == 0:
                    continue
                keep = np.r_[True, np.any(r[1:, [0, 2, 3]] != r[:-1, [0, 2, 3]], axis=1)]
                first_hit = np.full(int(r[:, 0].max()) + 1, np.iinfo(np.int64).max, dtype=np.int64)
                np.minimum.at(first_hit, r[:, 0], r[:, 4])
                u = r[keep]
                u_hit = first_hit[u[:, 0]]
                tie = cov.listing_tie_ranks(u[:, 0], u_hit, sequence, kmer_order)
                listing = np.lexsort((u[:, 3], u[:, 2], u[:, 0], tie, u_hit))
                u = u[listing]
                cs = np.maximum(0, u[:, 2] - ext) + offset
                ce = np.minimum(len(sequence), u[:, 3] + ext) + offset
                self.target_covers[i][j][rc].extend(zip(cs.tolist(), ce.tolist()))
                if not rc:
                    seen = set()
                    for pi in u[:, 0].tolist():
                        if pi not in seen:
                            seen.add(pi)
                            p = probes[pi]
                            self.probe_map_counts[p] = self.probe_map_counts.get(p, 0) + 1
            start += len(batch)

    # ------------------------------------------------------------------ aggregates
    def _compute_bp_covered_in_target_genomes(self):
        """coverage_analysis.py:255-280: length of the union of the covers."""
        self.bp_covered = {}
        for i, j, gnm, rc in self._iter_target_genomes():
            self.bp_covered.setdefault(i, {}).setdefault(j, {False: None, True: None})
            covers = self.target_covers[i][j][rc]
            total = 0
            if covers:
                a = np.array(covers, dtype=np.int64)
                a = a[a[:, 1] > a[:, 0]]
                if len(a):
                    a = a[np.argsort(a[:, 0], kind='stable')]
                    run_max = np.maximum.accumulate(a[:, 1])
                    new = np.r_[True, a[1:, 0] > run_max[:-1]]
                    gid = np.cumsum(new) - 1
                    ends = np.zeros(int(gid[-1]) + 1, dtype=np.int64)
                    np.maximum.at(ends, gid, a[:, 1])
                    total = int((ends - a[new, 0]).sum())
            self.bp_covered[i][j][rc] = total

    def _compute_average_coverage_in_target_genomes(self):
        """coverage_analysis.py:282-322."""
        self.average_coverage = {}
        for i, j, gnm, rc in self._iter_target_genomes():
            self.average_coverage.setdefault(i, {}).setdefault(j, {False: None, True: None})
            total_covered = sum(c[1] - c[0] for c in self.target_covers[i][j][rc])
            self.average_coverage[i][j][rc] = (float(total_covered) / gnm.size(False),
                                               float(total_covered) / gnm.size(True))

    def _compute_sliding_coverage_in_target_genomes(self, window_length, window_stride):
        """coverage_analysis.py:324-401: per-base depth (uint16, as the reference stores it), then the average
        over windows keyed by their middle position."""
        self.sliding_coverage = {}
        for i, j, gnm, rc in self._iter_target_genomes():
            self.sliding_coverage.setdefault(i, {}).setdefault(j, {False: None, True: None})
            covers = self.target_covers[i][j][rc]
            size = gnm.size(False)
            diff = np.zeros(size + 1, dtype=np.int64)
            if covers:
                a = np.array(covers, dtype=np.int64)
                np.add.at(diff, a[:, 0], 1)
                np.add.at(diff, a[:, 1], -1)
            probe_counts = np.cumsum(diff)[:size].astype('uint16')
            if covers:
                # the reference only fills positions between the first and the last endpoint (:362-379)
                lo, hi = int(a.min()), int(a.max())
                probe_counts[:lo] = 0
                probe_counts[hi:] = 0
            out = {}
            for window_start in np.arange(0, size, window_stride):
                window_end = window_start + window_length
                if window_end > size:
                    window_end = size
                    window_start = window_end - window_length
                middle = window_start + (window_length / 2)
                out[middle] = np.average(probe_counts[window_start:window_end])
            self.sliding_coverage[i][j][rc] = out

    def run(self, window_length=50, window_stride=25):
        self._find_covers_in_target_genomes()
        self._compute_bp_covered_in_target_genomes()
        self._compute_average_coverage_in_target_genomes()
        self._compute_sliding_coverage_in_target_genomes(window_length, window_stride)

    # ------------------------------------------------------------------ output (coverage_analysis.py:418-600)
    def write_data_matrix_as_tsv(self, fn):
        data = [["Genome", "Num bases covered", "Frac bases covered", "Frac bases covered over unambig",
                 "Average coverage/depth", "Average coverage/depth over unambig"]]
        for i, j, gnm, rc in self._iter_target_genomes():
            col_header = "%s, genome %d" % (self.target_genomes_names[i], j)
            if rc:
                col_header += " (rc)"
            bp_covered = self.bp_covered[i][j][rc]
            avg_all, avg_unambig = self.average_coverage[i][j][rc]
            data += [[col_header, bp_covered, float(bp_covered) / gnm.size(False),
                      float(bp_covered) / gnm.size(True), avg_all, avg_unambig]]
        with open(fn, 'w') as f:
            for row in data:
                f.write('\t'.join([str(entry) for entry in row]) + '\n')

    def _make_data_matrix_string(self):
        data = [["Genome", "Num bases covered\n[over unambig]", "Average coverage/depth\n[over unambig]"]]
        for i, j, gnm, rc in self._iter_target_genomes():
            col_header = "%s, genome %d" % (self.target_genomes_names[i], j)
            if rc:
                col_header += " (rc)"
            bp_covered = self.bp_covered[i][j][rc]
            frac_all = float(bp_covered) / gnm.size(False)
            frac_unambig = float(bp_covered) / gnm.size(True)
            prct_all = "<0.01%" if frac_all < 0.0001 else "{0:.2%}".format(frac_all)
            prct_unambig = "<0.01%" if frac_unambig < 0.0001 else "{0:.2%}".format(frac_unambig)
            bp_covered_str = "%d (%s) [%s]" % (bp_covered, prct_all, prct_unambig)
            avg_all, avg_unambig = self.average_coverage[i][j][rc]
            avg_all_str = "<0.01" if avg_all < 0.01 else "{0:.2f}".format(avg_all)
            avg_unambig_str = "<0.01" if avg_unambig < 0.01 else "{0:.2f}".format(avg_unambig)
            data += [[col_header, bp_covered_str, "%s [%s]" % (avg_all_str, avg_unambig_str)]]
        return data

    def print_analysis(self):
        print("NUMBER OF PROBES: %d" % len(self.probes))
        print()
        print(_table(self._make_data_matrix_string(), ["left", "right", "right"], header_underline=True))

    def write_sliding_window_coverage(self, fn):
        with open(fn, 'w') as f:
            for i, j, gnm, rc in self._iter_target_genomes():
                header = "%s, genome %d" % (self.target_genomes_names[i], j)
                if rc:
                    header += " (rc)"
                cov_d = self.sliding_coverage[i][j][rc]
                for pos in sorted(cov_d.keys()):
                    f.write('\t'.join([str(x) for x in [header, pos, cov_d[pos]]]) + '\n')

    def write_probe_map_counts(self, fn):
        with open(fn, 'w') as f:
            f.write('\t'.join(["Probe identifier", "Probe sequence", "Number sequences mapped to"]) + '\n')
            for p, count in self.probe_map_counts.items():
                f.write('\t'.join([str(x) for x in [p.identifier(), p.seq_str, count]]) + '\n')
