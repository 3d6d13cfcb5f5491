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


def iterate_fasta(fn, data_type='str', replace_degenerate=True, block_bytes=64 << 20):
    """Yield sequences one at a time (no upper-casing, no gap removal, as the reference's seq_io.iterate_fasta
    :178-232).  The file is read in blocks of `block_bytes` that are cut at a line end and scanned natively
    (`_fastpack.fasta_stream_block`); a sequence is held in memory only until the header that follows it.  A block
    with non-ASCII bytes hands the rest of the file to the Python line loop."""
    if data_type != 'str':
        raise ValueError("Unknown data_type " + data_type if data_type != 'np' else
                         "data_type 'np' (arrays of single characters) is not provided; sequences are str")
    if _fastpack is None or not hasattr(_fastpack, 'fasta_stream_block'):
        yield from _iterate_fasta_lines(_open(fn), [], replace_degenerate)
        return
    pieces = []                                            # runs of the sequence being read (str)
    with (gzip.open(fn, 'rb') if fn.endswith('.gz') else open(fn, 'rb')) as f:
        tail = b''
        while True:
            block = f.read(block_bytes)
            if not block:
                break
            block = tail + block
            # cut at the last line end whose successor is known (a '\r' at the very end may be half of a '\r\n')
            cut = block.rfind(b'\n') + 1
            if cut == 0:
                cut = block.rfind(b'\r', 0, len(block) - 1) + 1
            tail, block = block[cut:], block[:cut]
            if not block:
                continue                                   # one line longer than the block: keep reading
            if not block.isascii():
                import io
                rest = io.TextIOWrapper(io.BytesIO(block + tail + f.read()))
                yield from _iterate_fasta_lines(rest, pieces, replace_degenerate)
                return
            for item in _fastpack.fasta_stream_block(block, replace_degenerate):
                if item is None:
                    if pieces:
                        yield ''.join(pieces)
                    pieces = []
                else:
                    pieces.append(item.decode('ascii'))
        if tail:
            if not tail.isascii():
                import io
                yield from _iterate_fasta_lines(io.TextIOWrapper(io.BytesIO(tail)), pieces, replace_degenerate)
                return
            for item in _fastpack.fasta_stream_block(tail, replace_degenerate):
                if item is None:
                    if pieces:
                        yield ''.join(pieces)
                    pieces = []
                else:
                    pieces.append(item.decode('ascii'))
    if pieces:
        yield ''.join(pieces)


def _iterate_fasta_lines(f, cur, replace_degenerate=True):
    """The line loop of iterate_fasta on an open text file, continuing a sequence whose first runs are in `cur`."""
    cur = list(cur)
    with f:
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
