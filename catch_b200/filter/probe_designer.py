"""Orchestrator: candidates per grouping, then the filter chain (host plumbing).

Mirrors catch/filter/probe_designer.py:16-315: candidate generation per grouping (:249-268), filters applied in
order with input_is_grouped=True (:186-207), final probes = list(set(chain(*probes))) (:288); and, when
cluster_threshold is set, clustering of all sequences by MinHash sketches first (:78-184, on the device through
catch_b200/utils/cluster.py), the filters up to cluster_merge_after run per cluster, the rest on the merged
probes (:291-315).
"""
import itertools
import logging

from catch_b200 import genome
from catch_b200.filter import candidate_probes
from catch_b200.probe_batch import ProbeBatch
from catch_b200.utils import cluster

logger = logging.getLogger(__name__)


class ProbeDesigner:
    def __init__(self, genomes, filters, probe_length, probe_stride, allow_small_seqs=None,
                 seq_length_to_skip=None, cluster_threshold=None, cluster_merge_after=None,
                 cluster_method=None, cluster_fragment_length=None):
        self.genomes = genomes
        self.filters = filters
        self.probe_length = probe_length
        self.probe_stride = probe_stride
        self.allow_small_seqs = allow_small_seqs
        self.seq_length_to_skip = seq_length_to_skip
        self.cluster_threshold = cluster_threshold
        self.cluster_merge_after = cluster_merge_after
        self.cluster_method = cluster_method
        self.cluster_fragment_length = cluster_fragment_length

    def _cluster_genomes(self):
        """All sequences of all groupings (or their fragments) clustered by nucleotide similarity; one Genome per
        sequence, one grouping per cluster, largest cluster first (:78-184)."""
        if len(self.genomes) > 1:
            logger.warning("There are >1 groups of genomes in the input, but clustering these will override those "
                           "groupings; differential identification or other tasks that rely on group separation "
                           "may no longer work as intended")
        seqs = {}
        seq_idx = 0
        for genomes_from_group in self.genomes:
            for g in genomes_from_group:
                if self.cluster_fragment_length is not None:
                    g_seqs = g.break_into_fragments(self.cluster_fragment_length, include_full_end=True).seqs
                else:
                    g_seqs = g.seqs
                for s in g_seqs:
                    if self.seq_length_to_skip is not None and len(s) <= self.seq_length_to_skip:
                        continue
                    seqs[seq_idx] = s
                    seq_idx += 1
        method = self.cluster_method
        if method == 'choose':                                  # :116-160
            method = 'simple'
            if self.cluster_fragment_length is not None:
                num_sequences = sum(len(g.seqs) for gs in self.genomes for g in gs)
                total_seq_len = sum(g.size() for gs in self.genomes for g in gs)
                if num_sequences > 1 and total_seq_len / num_sequences > self.cluster_fragment_length:
                    method = 'hierarchical'
        logger.info("Clustering %d sequences using MinHash signatures, at an average nucleotide dissimilarity "
                    "threshold of %f", seq_idx, self.cluster_threshold)
        clusters = cluster.cluster_with_minhash_signatures(seqs, threshold=self.cluster_threshold,
                                                           cluster_method=method)
        logger.info("Found %d clusters with sizes: %s", len(clusters), [len(c) for c in clusters])
        return [[genome.Genome.from_one_seq(seqs[i]) for i in clust] for clust in clusters]

    def design(self):
        if self.cluster_threshold is None:
            self._candidates, probes = self._design_for_genomes(self.genomes, self.filters)
            self.final_probes = list(set(itertools.chain(*probes)))
            return
        assert self.cluster_merge_after is not None                      # :291-315
        assert self.cluster_merge_after in self.filters
        merge_idx = self.filters.index(self.cluster_merge_after) + 1
        clustered_genomes = self._cluster_genomes()
        self._candidates, probes_by_cluster = self._design_for_genomes(clustered_genomes, self.filters[:merge_idx])
        probes = list(set(itertools.chain(*probes_by_cluster)))
        for f in self.filters[merge_idx:]:
            logger.info("Starting filter %s", f.__class__.__name__)
            probes = f.filter(probes, clustered_genomes, input_is_grouped=False)
        self.final_probes = probes

    def _design_for_genomes(self, genomes, filters):
        candidates = []
        for genomes_from_group in genomes:
            group = self._candidates_as_batch(genomes_from_group)
            if group is None:                   # small sequences in play: the per-object path knows those rules
                group = []
                for g in genomes_from_group:
                    group += candidate_probes.make_candidate_probes_from_sequences(
                        g.seqs, probe_length=self.probe_length, probe_stride=self.probe_stride,
                        allow_small_seqs=self.allow_small_seqs, seq_length_to_skip=self.seq_length_to_skip)
            if not len(group):
                logger.warning("There are no candidate probes for a grouping of genomes")
            candidates.append(group)
        probes = candidates
        for f in filters:
            logger.info("Starting filter %s", f.__class__.__name__)
            probes = f.filter(probes, genomes, input_is_grouped=True)
        return candidates, probes

    @property
    def candidate_probes(self):
        """All candidates as Probe objects (probe_designer.py:268); materialised on demand only."""
        return list(itertools.chain(*self._candidates))

    def _candidates_as_batch(self, genomes_from_group):
        """The candidates of one grouping as one buffer (catch_b200/probe_batch.py), or None when a sequence is
        shorter than the probe length (then --small-seq-min / the error of candidate_probes.py:53-70 applies)."""
        seqs = []
        for g in genomes_from_group:
            if not isinstance(g.seqs, list) or len(g.seqs) == 0 or not all(isinstance(s, str) for s in g.seqs):
                return None
            seqs.extend(g.seqs)
        # all sequences of the grouping in one call (genome order, then sequence order, as :249-256 concatenates them)
        out = ProbeBatch.from_sequences(seqs, self.probe_length, self.probe_stride,
                                        seq_length_to_skip=self.seq_length_to_skip)
        if out is None:
            return None
        return out if len(out) else []
