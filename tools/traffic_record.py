#!/usr/bin/env python
"""Record the DRAM traffic of `ncu --set full` captures for bench.py's `roofline.traffic`.

    python tools/traffic_record.py 'scan_kernel=gpurun_out/prof_scan_r02b.ncu-rep:catch_b200/csrc/coverage.cu' ...

Writes profiles/traffic_r02.json: per kernel label the dram__bytes_read/write sums of ONE launch and the sha256 of the
kernel's source file at the time of the capture; bench.py quotes the number only while that hash still matches."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    kernels = []
    for arg in sys.argv[1:]:
        label, rest = arg.split('=', 1)
        rep, source = rest.split(':', 1)
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]

        def get(name):
            i = hdr.index(name)
            return float(vals[i].replace(',', '')) * UNIT[units[i]]
        kernels.append({
            'label': label, 'kernel': vals[hdr.index('Kernel Name')], 'capture': 'profiles/' + os.path.basename(rep).replace('.ncu-rep', '_metrics.csv'),
            'dram_read_bytes': get('dram__bytes_read.sum'), 'dram_write_bytes': get('dram__bytes_write.sum'),
            'duration_ms': float(vals[hdr.index('gpu__time_duration.sum')].replace(',', '')) *
                           {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[units[hdr.index('gpu__time_duration.sum')]],
            'source': source, 'source_sha256': hashlib.sha256(open(os.path.join(ROOT, source), 'rb').read()).hexdigest(),
        })
    with open(os.path.join(ROOT, 'profiles', 'traffic_r02.json'), 'w') as f:
        json.dump({'workload': 'config 2 (Zika-scale), python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline',
                   'kernels': kernels}, f, indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == '__main__':
    main()
