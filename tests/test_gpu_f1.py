"""the consumers of the stage-A scan next to the set cover -- AdapterFilter and the
coverage Analyzer -- on the device scan, against fixtures recorded by RUNNING THE REFERENCE
(tests/golden/make_f1_golden.py: every call the reference's own test_adapter_filter.py /
test_coverage_analysis.py make, plus seeded random cases).  Their output order depends on Python's
string hash seed (as the reference's does), so the replay runs in a child interpreter under
PYTHONHASHSEED=0, the seed the fixtures were recorded with.  Needs a B200."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_REPLAY = r"""
import json, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from tests import golden_io, helpers
from catch_b200 import _lib, probe
from catch_b200 import coverage_analysis as ca
from catch_b200.filter import adapter_filter as af

ctx = _lib.default_context()
gold = golden_io.load('f1_reference.json.gz')


def set_state(st):
    np.random.set_state(('MT19937', np.array(st[0], dtype=np.uint32), st[1], 0, 0.0))


bad = {'adapter': [], 'analyzer': []}
for n, r in enumerate(gold['adapter']):
    f = af.AdapterFilter(tuple(r['adapter_a']), tuple(r['adapter_b']), mismatches=r['mismatches'],
                         lcf_thres=r['lcf_thres'], kmer_probe_map_k=r['k'])
    f._ctx = ctx
    set_state(r['np_state'])
    out = f.filter([probe.Probe.from_str(s) for s in r['probes']], helpers.to_genomes(r['genomes']))
    if [p.seq_str for p in out] != r['out']:
        bad['adapter'].append(n)


def flat(d, conv):
    return [[i, j, int(rc), conv(v)] for i in sorted(d) for j in sorted(d[i]) for rc, v in sorted(d[i][j].items())
            if v is not None]


for n, r in enumerate(gold['analyzer']):
    an = ca.Analyzer([probe.Probe.from_str(s) for s in r['probes']], r['mismatches'], r['lcf_thres'],
                     helpers.to_genomes(r['genomes']), r['names'], cover_extension=r['cover_extension'],
                     kmer_probe_map_k=r['k'], rc_too=r['rc_too'])
    an._ctx = ctx
    set_state(r['np_state'])
    an.run(*r['window'])
    got = dict(
        target_covers=flat(an.target_covers, lambda v: [[int(a), int(b)] for a, b in v]),
        bp_covered=flat(an.bp_covered, int),
        average_coverage=flat(an.average_coverage, lambda v: [float(v[0]), float(v[1])]),
        sliding_coverage=flat(an.sliding_coverage, lambda v: [[float(k), float(x)] for k, x in sorted(v.items())]),
        probe_map_counts=[[p.seq_str, int(c)] for p, c in an.probe_map_counts.items()],
        table=an._make_data_matrix_string())
    wrong = [k for k in got if got[k] != r[k]]
    if wrong:
        bad['analyzer'].append([n, wrong])
print(json.dumps({'n_adapter': len(gold['adapter']), 'n_analyzer': len(gold['analyzer']), 'bad': bad}))
"""


def test_adapter_filter_and_analyzer_against_the_reference():
    env = dict(os.environ, PYTHONHASHSEED='0')
    res = subprocess.run([sys.executable, '-c', _REPLAY, ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out['n_adapter'] >= 20 and out['n_analyzer'] >= 20
    assert out['bad'] == {'adapter': [], 'analyzer': []}, out


def test_scan_records_match_reference_scans(ctx):
    """cb_coverage_records (the unmerged ranges of the scan) against every find_probe_covers_in_sequence vector
    of the reference's test_probe.py: merged on the host the way probe.py:1262-1270 does for
    merge_overlapping=True (interval.merge_overlapping) and as sorted distinct ranges for False."""
    import numpy as np
    from catch_b200 import coverage as cov
    from tests import golden_io
    ref = golden_io.load('reference_tests.json.gz')
    n = 0
    for r in ref['scan']:
        group = cov.PackedGroup(ctx, r['probes'], [[r['seq']]])
        off = np.zeros(len(r['seeds']) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in r['seeds']])
        flat = [int(x) for sl in r['seeds'] for x in sl]
        pos = np.array(flat if flat else [0], dtype=np.int32)
        rec, _ = ctx.coverage_records(group.probes, group.targets, r['m'], r['lcf'], r['island'], r['k'], off, pos)
        group.free()
        per_probe = {}
        for p, q, s, e, h in sorted(set(map(tuple, rec.tolist()))):
            assert q == 0 and 0 <= h <= len(r['seq']) - r['k']
            per_probe.setdefault(r['probes'][p], []).append([s, e])
        got = {}
        for k, ranges in per_probe.items():
            ranges = sorted(map(list, set(map(tuple, ranges))))
            if r['merge']:
                merged = []
                for s, e in ranges:
                    if merged and s <= merged[-1][1]:
                        merged[-1][1] = max(merged[-1][1], e)
                    else:
                        merged.append([s, e])
                ranges = merged
            got[k] = ranges
        assert got == r['out'], (r['probes'], r['seq'][:80])
        n += 1
    assert n >= 100
