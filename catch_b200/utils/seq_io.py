"""FASTA wire format (host plumbing), byte-compatible with the reference's catch/utils/seq_io.py:
reading upper-cases, maps degenerate bases [YRWSMKBDHV] to N and drops '-' (:130,149-154);
writing emits '>probe_<identifier>' / sequence pairs (:235-252)."""
import gzip
import re
from collections import OrderedDict

from catch_b200 import genome

_DEGENERATE = re.compile('[YRWSMKBDHV]')


def _open(fn):
    return gzip.open(fn, 'rt') if fn.endswith('.gz') else open(fn, 'r')


def read_fasta(fn, replace_degenerate=True, skip_gaps=True, make_uppercase=True):
    """Ordered mapping name -> sequence (str)."""
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
    with open(out_fn, 'w') as f:
        for p in probes:
            f.write('>%s\n' % (p.header if p.header else 'probe_%s' % p.identifier()))
            f.write(p.seq_str + '\n')
