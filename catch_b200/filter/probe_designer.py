"""Orchestrator: candidates per grouping, then the filter chain (host plumbing).

Mirrors the non-clustering path of catch/filter/probe_designer.py:16-315: candidate generation per
grouping (:249-268), filters applied in order with input_is_grouped=True (:186-207), final probes =
list(set(chain(*probes))) (:288).  Genome clustering (--cluster-and-design-separately) is outside the
hot path and not provided.
"""
import itertools
import logging

from catch_b200.filter import candidate_probes
from catch_b200.probe_batch import ProbeBatch

logger = logging.getLogger(__name__)


class ProbeDesigner:
    def __init__(self, genomes, filters, probe_length, probe_stride, allow_small_seqs=None,
                 seq_length_to_skip=None, cluster_threshold=None, cluster_merge_after=None,
                 cluster_method=None, cluster_fragment_length=None):
        if cluster_threshold is not None:
            raise NotImplementedError("genome clustering (catch/utils/cluster.py) is outside the "
                                      "accelerated hot path; run without --cluster-and-design-separately")
        self.genomes = genomes
        self.filters = filters
        self.probe_length = probe_length
        self.probe_stride = probe_stride
        self.allow_small_seqs = allow_small_seqs
        self.seq_length_to_skip = seq_length_to_skip

    def design(self):
        candidates = []
        for genomes_from_group in self.genomes:
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
        for f in self.filters:
            logger.info("Starting filter %s", f.__class__.__name__)
            probes = f.filter(probes, self.genomes, input_is_grouped=True)
        self._candidates = candidates
        self.final_probes = list(set(itertools.chain(*probes)))

    @property
    def candidate_probes(self):
        """All candidates as Probe objects (probe_designer.py:268); materialised on demand only."""
        return list(itertools.chain(*self._candidates))

    def _candidates_as_batch(self, genomes_from_group):
        """The candidates of one grouping as one buffer (catch_b200/probe_batch.py), or None when a sequence is
        shorter than the probe length (then --small-seq-min / the error of candidate_probes.py:53-70 applies)."""
        batches = []
        for g in genomes_from_group:
            if not isinstance(g.seqs, list) or len(g.seqs) == 0 or not all(isinstance(s, str) for s in g.seqs):
                return None
            b = ProbeBatch.from_sequences(g.seqs, self.probe_length, self.probe_stride,
                                          seq_length_to_skip=self.seq_length_to_skip)
            if b is None:
                return None
            batches.append(b)
        out = ProbeBatch.concat(batches)
        return out if out is not None else []
