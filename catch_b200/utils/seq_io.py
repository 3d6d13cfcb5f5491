"""FASTA wire format (host plumbing), byte-compatible with the reference's catch/utils/seq_io.py:
reading upper-cases, maps degenerate bases [YRWSMKBDHV] to N and drops '-' (:130,149-154);
writing emits '>probe_<identifier>' / sequence pairs (:235-252).

Fast path: an ASCII file is parsed by one native pass over its bytes
(`_fastpack.parse_fasta`, csrc/fastpack.c) instead of a Python loop over its lines; files with
non-ASCII bytes, and builds without the helper, take the line loop below, which is the same rule
set written in Python.  tools/fasta_bench.py times both against the reference's reader."""
import gzip
import re
from collections import OrderedDict

from catch_b200 import genome

try:
    from catch_b200 import _fastpack
except ImportError:                                    # helper not built: the line loop does the same
    _fastpack = None

_DEGENERATE = re.compile('[YRWSMKBDHV]')


def _open(fn):
    return gzip.open(fn, 'rt') if fn.endswith('.gz') else open(fn, 'r')


def read_fasta(fn, data_type='str', replace_degenerate=True, skip_gaps=True, make_uppercase=True):
    """Ordered mapping name -> sequence (str), as seq_io.read_fasta (:104-175)."""
    if data_type != 'str':
        raise ValueError("Unknown data_type " + data_type if data_type != 'np' else
                         "data_type 'np' (arrays of single characters) is not provided; sequences are str")
    if _fastpack is not None and hasattr(_fastpack, 'parse_fasta'):
        with (gzip.open(fn, 'rb') if fn.endswith('.gz') else open(fn, 'rb')) as f:
            data = f.read()
        if data.isascii():
            names, seqs = _fastpack.parse_fasta(data, make_uppercase, replace_degenerate, skip_gaps)
            m = OrderedDict()
            for name, seq in zip(names, seqs):
                m[name] = seq                          # a repeated header replaces the entry in place (:145-146)
            return m
    return _read_fasta_lines(fn, replace_degenerate, skip_gaps, make_uppercase)


def _read_fasta_lines(fn, replace_degenerate=True, skip_gaps=True, make_uppercase=True):
    chunks = OrderedDict()
    name = ""
    with _open(fn) as f:
        for line in f:
            line = line.rstrip()
            if not line:
                name = ""
                continue
            if name == "":
                assert line.startswith('>')
            if line.startswith('>'):
                name = line[1:]
                chunks[name] = []
                continue
            if make_uppercase:
                line = line.upper()
            if replace_degenerate:
                line = _DEGENERATE.sub('N', line)
            if skip_gaps:
                line = line.replace('-', '')
            chunks[name].append(line)
    return OrderedDict((k, ''.join(v)) for k, v in chunks.items())


def iterate_fasta(fn, replace_degenerate=True):
    """Yield sequences one at a time (no upper-casing, as reference :178-232)."""
    cur = []
    with _open(fn) as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            if line.startswith('>'):
                if cur:
                    yield ''.join(cur)
                cur = []
            else:
                if replace_degenerate:
                    line = _DEGENERATE.sub('N', line)
                cur.append(line)
    if cur:
        yield ''.join(cur)


def read_genomes_from_fasta(fn):
    """One Genome per FASTA record (:85-101)."""
    return [genome.Genome.from_one_seq(s) for s in read_fasta(fn).values()]


def write_probe_fasta(probes, out_fn):
    """'>header' or '>probe_<identifier>' and the sequence, one pair per probe (:235-252); one write call."""
    parts = []
    for p in probes:
        parts.append('>%s\n%s\n' % (p.header if p.header else 'probe_%s' % p.identifier(), p.seq_str))
    with open(out_fn, 'w') as f:
        f.write(''.join(parts))
