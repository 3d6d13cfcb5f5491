import sys, time, os, subprocess
sys.path.insert(0, os.getcwd())
if len(sys.argv) > 1:
    import numpy as np
    from catch_b200 import _fastpack, probe
    rng = np.random.default_rng(0)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    out = []
    for mb in (4, 7, 13, 26, 52, 133):
        n = mb * 1000000 // 100
        flat = letters[rng.integers(0, 4, (n, 100), dtype=np.uint8)].tobytes().decode()
        probes = [probe.Probe(flat[i * 100:(i + 1) * 100]) for i in range(n)]
        best = 1e9
        for _ in range(7):
            t = time.perf_counter(); d, l = _fastpack.gather(probes, 'seq_str'); best = min(best, time.perf_counter() - t)
        out.append('%dMB %.2f' % (mb, best * 1e3))
    print('par>=%sMB:' % sys.argv[1], ' '.join(out), flush=True)
else:
    for thr in ('4', '8', '32', '100000'):
        subprocess.run([sys.executable, __file__, thr], env=dict(os.environ, CB_GATHER_PAR_MB=thr))
