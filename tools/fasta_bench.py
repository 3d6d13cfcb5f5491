#!/usr/bin/env python
"""FASTA reader throughput: catch_b200.utils.seq_io.read_fasta (native one-pass parser) against the
Python line loop with the same rules (`_read_fasta_lines`) and, when /root/reference is present, the reference's
seq_io.read_fasta, on one synthetic file (70-column lines, lower-case and degenerate bases sprinkled in).
Host-only; prints one JSON line.

    python tools/fasta_bench.py [--mb 64]
"""
import argparse
import json
import logging
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from catch_b200.utils import seq_io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mb', type=int, default=64)
    ap.add_argument('--repeat', type=int, default=3)
    args = ap.parse_args()
    logging.disable(logging.CRITICAL)
    rng = np.random.default_rng(0)
    n_rec = max(1, args.mb * 1000000 // 13588)
    letters = np.frombuffer(b'ACGT' * 60 + b'acgtNRYK-', dtype=np.uint8)
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, 'in.fasta')
        with open(fn, 'wb') as f:
            for i in range(n_rec):
                s = letters[rng.integers(0, len(letters), 13588)].tobytes()
                f.write(b'>g%d\n' % i)
                f.write(b'\n'.join(s[j:j + 70] for j in range(0, len(s), 70)) + b'\n')
        size = os.path.getsize(fn)
        out = dict(file_mb=round(size / 1e6, 1), records=n_rec)
        readers = [('native', seq_io.read_fasta), ('python_lines', seq_io._read_fasta_lines)]
        if os.path.isdir('/root/reference'):
            sys.path.insert(0, '/root/reference')
            from catch.utils import seq_io as rseq_io
            readers.append(('reference', rseq_io.read_fasta))
        results = {}
        for name, fn_read in readers:
            best = None
            for _ in range(args.repeat if name == 'native' else 1):
                t0 = time.perf_counter()
                m = fn_read(fn)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            results[name] = m
            out[name + '_s'] = round(best, 3)
            out[name + '_mb_per_s'] = round(size / 1e6 / best, 1)
        for name in results:
            assert list(results[name].items()) == list(results['native'].items()), name
        out['identical'] = True
    print(json.dumps(out))


if __name__ == '__main__':
    main()
