"""Candidate probes by tiling (host plumbing feeding the hot path).

Behaviour of the reference's catch/filter/candidate_probes.py:21-182: starts 0, stride, ... while the
probe fits (:97-100); one extra probe flush with the end when len % stride != 0 (:102-106); probes
containing a run of >= min_n_string_length N are dropped and probes flanking every such run are
added (:47, :112-122).
"""
import logging
import re

from catch_b200 import probe

logger = logging.getLogger(__name__)


def make_candidate_probes_from_sequence(seq, probe_length, probe_stride, min_n_string_length=2,
                                        allow_small_seqs=None):
    if not isinstance(seq, str):
        seq = ''.join(seq)
    n_run = re.compile('(N{%d,})' % min_n_string_length)
    if len(seq) < probe_length:
        if not allow_small_seqs:
            raise ValueError("An input sequence is smaller than the probe length (%d); try setting "
                             "--small-seq-skip" % probe_length)
        if len(seq) < allow_small_seqs:
            raise ValueError("Allowing sequences smaller than the probe length (%d), but input sequence "
                             "is smaller than minimum allowed length" % probe_length)
        if n_run.search(seq):
            raise Exception("Only possible probe from input sequence has too long a stretch of N's")
        return [probe.Probe.from_str(seq)]

    out = []

    def take(start, end, flanking=False):
        sub = seq[start:end]
        if n_run.search(sub) is None:
            p = probe.Probe.from_str(sub)
            p.is_flanking_n_string = flanking
            out.append(p)

    start = 0
    while start + probe_length <= len(seq):
        take(start, start + probe_length)
        start += probe_stride
    if len(seq) % probe_stride != 0:
        take(len(seq) - probe_length, len(seq))
    for m in n_run.finditer(seq):
        if m.start() - probe_length >= 0:
            take(m.start() - probe_length, m.start(), True)
        if m.end() + probe_length <= len(seq):
            take(m.end(), m.end() + probe_length, True)
    return out


def make_candidate_probes_from_sequences(seqs, probe_length, probe_stride, min_n_string_length=2,
                                         allow_small_seqs=None, seq_length_to_skip=None):
    if not isinstance(seqs, list):
        raise TypeError("seqs must be a list of sequences")
    if len(seqs) == 0:
        raise ValueError("seqs must have at least one sequence")
    for s in seqs:
        if not isinstance(s, str):
            raise TypeError("seqs must be a list of Python strings")
    out = []
    for s in seqs:
        if seq_length_to_skip is not None and len(s) <= seq_length_to_skip:
            logger.info("Not designing candidate probes for a sequence with length %d, since it is <= %d",
                        len(s), seq_length_to_skip)
            continue
        out += make_candidate_probes_from_sequence(s, probe_length, probe_stride, min_n_string_length,
                                                   allow_small_seqs)
    return out
