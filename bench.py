#!/usr/bin/env python
"""Benchmark of the SetCoverFilter hot path (BASELINE.json metric: candidate-probe x target-bp / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload zika|plumbing|influenza|vall|sweep]
                    [--shard probes|groups] [--impl reference] [--no-extras] [--no-cpu-baseline]

A step is one pass of the hot path (stage A coverage + stage B greedy set cover) over one batch
of synthetic genomes.  Two numbers per run:
  value : inputs already packed and resident in HBM when the timed region starts (device
          timeline, CUDA events on the library's stream);
  e2e   : the same metric through the plugin call a user makes, SetCoverFilter.filter(), with
          host Probe/Genome objects; host packing, host->device copies, both stages and the
          device->host read of the selection are all inside the timed region.
With N > 1 (torchrun, one process per GPU) the default is STRONG scaling on the same single
grouping: the candidate probes are sharded over the GPUs, stage A runs on each shard, and stage B
runs with sharded gains and one exchange per greedy round through peer-mapped memory
(csrc/rounds.cu).  `--shard groups` gives one independent grouping per GPU instead (weak scaling).

The default line (workload zika = BASELINE config 2) also carries, unless --no-extras:
  like_for_like : SetCoverFilter.filter() on the SAME 60-genome sample the CPU leg runs, with the
                  selection compared to the oracle's (bit-exact, order included);
  configs       : compact results for BASELINE configs 3 (influenza shape, MinHash near-duplicate
                  filter on), 4 (V-All shape, a stated number of taxa) and 5 (m x l sweep);
  with N > 1: vall_groups, the V-All-shape e2e with the SAME taxa sharded over the N GPUs, and influenza_shard, the
  config-3 set cover (one grouping, 61 k probes x 68 Mbp) with its probes sharded over the N GPUs.
The CPU oracle (oracle/) is executed only for `cpu_baseline`, `like_for_like` and `--impl reference`.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import helpers  # noqa: E402  (synthetic generators of SURVEY.md section 8d)

RNG_SEED = 7
METRIC = 'candidate-probe x target-bp / s through SetCoverFilter'

WORKLOADS = {
    # single-grouping set-cover workloads: n_genomes, genome_len, divergence, generator seed, probe length / stride
    'plumbing': dict(n_genomes=20, length=5000, div=0.03, seed=1, pl=75, ps=50,
                     scf=dict(mismatches=0, lcf_thres=75, cover_extension=0),
                     desc='config 1: 20 x 5 kb, -pl 75 -m 0 -e 0'),
    'zika': dict(n_genomes=500, length=11000, div=0.03, seed=2, pl=75, ps=50,
                 scf=dict(mismatches=2, lcf_thres=60, cover_extension=50),
                 desc='config 2 (Zika-scale): 500 x 11 kb, -pl 75 -m 2 -l 60 -e 50'),
}
OTHER_WORKLOADS = ('influenza', 'vall', 'sweep')


def make_workload(name, n_genomes=None, n_groups=1):
    """n_groups independent groupings of the named shape (generator seeds seed, seed+1, ...)."""
    w = dict(WORKLOADS[name])
    if n_genomes is not None:
        w['n_genomes'] = n_genomes
    w['groups_seqs'], w['groups_cands'] = [], []
    pairs = 0
    for g in range(n_groups):
        seqs = helpers.synthetic_genomes(w['n_genomes'], w['length'], w['div'], w['seed'] + g)
        cands = helpers.tile_candidates(seqs, w['pl'], w['ps'])
        cands = list(dict.fromkeys(cands))          # DuplicateFilter upstream of SetCoverFilter
        w['groups_seqs'].append(seqs)
        w['groups_cands'].append(cands)
        pairs += len(cands) * sum(len(s) for s in seqs)
    w['seqs'], w['cands'] = w['groups_seqs'][0], w['groups_cands'][0]
    w['pairs'] = pairs
    return w


def sample_desc(w, name):
    return 'first %d genomes of %s (P=%d, T=%d bp)' % (w['n_genomes'], name, len(w['cands']), sum(map(len, w['seqs'])))


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def captured_traffic(kernel_label):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set
    full` capture (profiles/traffic_r02.json, written by tools/ncu_summary.py --traffic).  DRAM bytes cannot be
    measured inside the run; the capture is only quoted while the kernel source it was taken from is the source in
    the tree (sha256 of the .cu file recorded with it) and the workload is config 2 -- otherwise null, never a stale
    constant."""
    import hashlib
    path = os.path.join(ROOT, 'profiles', 'traffic_r02.json')
    try:
        rec = json.load(open(path))
        for k in rec['kernels']:
            if kernel_label.startswith(k['label']):
                src = os.path.join(ROOT, k['source'])
                if hashlib.sha256(open(src, 'rb').read()).hexdigest() != k['source_sha256']:
                    return None, '%s changed since the ncu capture in %s was taken' % (k['source'], k['capture'])
                return k['dram_read_bytes'] + k['dram_write_bytes'], \
                    'ncu --set full capture %s (same kernel source, config 2), per launch' % k['capture']
    except (OSError, KeyError, ValueError):
        pass
    return None, 'no ncu capture of this kernel recorded in profiles/traffic_r02.json'


def host_threads():
    """Host threads the CPU leg may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU leg
    is a reported baseline on 'all the host threads it can use', so it sizes itself from the machine."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------
def oracle_scf(w, threads):
    """One SetCoverFilter pass of the CPU restatement of the reference on workload w; returns
    (seconds, selected indices in the reference's output order)."""
    from oracle import oracle as O
    np.random.seed(RNG_SEED)
    random.seed(RNG_SEED)
    t0 = time.perf_counter()
    sel = O.set_cover_filter([w['cands']], [[[s] for s in w['seqs']]], w['scf']['mismatches'],
                             w['scf']['lcf_thres'], 0, 1.0, w['scf']['cover_extension'], 20, n_threads=threads)
    return time.perf_counter() - t0, sel[0]


def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU restatement of the reference path (oracle port -- the reference
    itself is Python and does not travel to the GPU box) on the host cores, bounded sample.  Rank 0
    alone runs it, with every host thread of the box, whatever N is."""
    if rank != 0:
        return
    name = args.workload if args.workload in WORKLOADS else 'zika'
    w = make_workload(name, n_genomes=args.cpu_sample_genomes)
    threads = host_threads()
    for _ in range(args.warmup):
        oracle_scf(w, threads)
    times = [oracle_scf(w, threads)[0] for _ in range(args.steps)]
    t = sum(times) / len(times)
    v = w['pairs'] / t
    sample = sample_desc(w, name)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC,
        'value': v, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[name]['desc'], 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def cpu_and_like_for_like(args, ctx, name):
    """The CPU leg and the GPU path on the SAME bounded sample: the oracle port on all
    host threads, SetCoverFilter.filter() on the device, the two selections compared element by element."""
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    w = make_workload(name, n_genomes=args.cpu_sample_genomes)
    threads = host_threads()
    t_cpu, want = oracle_scf(w, threads)
    sample = sample_desc(w, name)
    cpu = {'value': w['pairs'] / t_cpu, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
           'sample': '%s, %.1f s; stage A on %d threads, stage B on 1 (as the reference: one process per grouping)'
                     % (sample, t_cpu, threads)}
    genomes = helpers.to_genomes([[[s] for s in w['seqs']]])
    probes = [[probe.Probe.from_str(s) for s in w['cands']]]
    scf = SetCoverFilter(**w['scf'])
    scf._ctx = ctx
    os.environ['CB_SHARD'] = 'groups'
    times, got = [], None
    for i in range(5):
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        ctx.flush_l2()
        t0 = time.perf_counter()
        out = scf.filter(probes, genomes, input_is_grouped=True)
        times.append(time.perf_counter() - t0)
        ids = {id(p): i for i, p in enumerate(probes[0])}
        got = [ids[id(p)] for p in out[0]]
    t_gpu = float(np.mean(times[2:]))
    lfl = {'sample': sample, 'gpu_ms': t_gpu * 1e3, 'cpu_ms': t_cpu * 1e3, 'ratio': t_cpu / t_gpu,
           'gpu_pairs_per_s': w['pairs'] / t_gpu, 'cpu_pairs_per_s': w['pairs'] / t_cpu, 'cpu_cores': threads,
           'selected': len(got), 'identical': got == list(want),
           'what': 'SetCoverFilter.filter() end to end on host objects vs the CPU port of the reference on the same '
                   'input and RNG seed; identical = same probes in the same output order'}
    return cpu, lfl


# ------------------------------------------------------------------------------------------
def config3_influenza(ctx, n_genomes, reps=3):
    """BASELINE config 3: influenza shape, MinHash near-duplicate filter, then SetCoverFilter on what it keeps
    (-m 5 -l 30 -e 50, pl 100).  Returns (compact result, kept probes, genomes, T)."""
    from catch_b200 import probe
    from catch_b200.filter.near_duplicate_filter import NearDuplicateFilterWithMinHash
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    gens = helpers.synthetic_influenza(n_genomes, seed=3)
    groups = [[[seg] for g in gens for seg in g]]
    genomes = helpers.to_genomes(groups)
    cands = helpers.tile_candidates([s for g in groups[0] for s in g], 100, 50)
    T = sum(len(s) for g in groups[0] for s in g)
    raw = [[probe.Probe.from_str(s) for s in cands]]
    ndf_wall, ndf_dev, kept, st = [], [], None, None
    for _ in range(reps):
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        ndf = NearDuplicateFilterWithMinHash(0.6)
        ndf._ctx = ctx
        t = time.perf_counter()
        kept = ndf.filter(raw, genomes, input_is_grouped=True)
        ndf_wall.append(time.perf_counter() - t)
        st = ndf.last_stats
        ndf_dev.append(st['ms_total'])
    # the same filter on the candidates as ONE buffer (ProbeBatch: what the tiling of design.py hands to the filter
    # chain) instead of 1.3 M Probe objects
    from catch_b200.probe_batch import ProbeBatch
    batch = [ProbeBatch(np.frombuffer(''.join(cands).encode(), dtype=np.uint8).reshape(len(cands), 100))]
    ndf_wall_b, kept_b = [], None
    for _ in range(reps):
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        ndf = NearDuplicateFilterWithMinHash(0.6)
        ndf._ctx = ctx
        t = time.perf_counter()
        kept_b = ndf.filter(batch, genomes, input_is_grouped=True)
        ndf_wall_b.append(time.perf_counter() - t)
    same_b = [p.seq_str for p in kept_b[0]] == [p.seq_str for p in kept[0]]
    del batch, kept_b
    scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=50)
    scf._ctx = ctx
    os.environ['CB_SHARD'] = 'groups'
    scf_wall, n_sel = [], 0
    for _ in range(reps):
        np.random.seed(RNG_SEED)
        ctx.flush_l2()
        t = time.perf_counter()
        out = scf.filter(kept, genomes, input_is_grouped=True)
        scf_wall.append(time.perf_counter() - t)
        n_sel = len(out[0])
    s = scf.last_stats[0]
    P = len(kept[0])
    t_ndf, t_scf = min(ndf_wall[1:] or ndf_wall), min(scf_wall[1:] or scf_wall)
    dev_ms = s['coverage']['ms_total'] + s['setcover']['ms_total']
    res = {
        'workload': 'config 3 (influenza shape): %d genomes x 8 segments (T=%d bp), pl 100 ps 50, MinHash near-duplicate '
                    'filter 0.6, then -m 5 -l 30 -e 50' % (n_genomes, T),
        'P_raw': len(cands), 'P_distinct': int(st['n_distinct']), 'P_after_ndf': P,
        'ndf': {'wall_ms': t_ndf * 1e3, 'wall_batch_ms': min(ndf_wall_b[1:] or ndf_wall_b) * 1e3, 'batch_output_identical': same_b,
                'device_ms': min(ndf_dev), 'probes_per_s_e2e': len(cands) / t_ndf,
                'probes_per_s_device': len(cands) / (min(ndf_dev) / 1e3), 'decision_rounds': int(st['n_picks']),
                'device_split_ms': {'grouping': st['ms_pack'], 'signatures': st['ms_seed_index'], 'decisions': st['ms_greedy']}},
        'scf': {'e2e_ms': t_scf * 1e3, 'device_ms': dev_ms, 'pairs_per_s_e2e': P * T / t_scf,
                'pairs_per_s_device': P * T / (dev_ms / 1e3), 'selected': n_sel,
                'scan_ms': s['coverage']['ms_scan_emit'], 'merge_ms': s['coverage']['ms_merge'],
                'greedy_ms': s['setcover']['ms_greedy'], 'rounds': int(s['setcover']['reserved'][5])},
    }
    return res, kept, genomes, T


def config3_sharded(ctx, n_genomes, dist, local_rank, reps=3):
    """BASELINE config 3 (influenza shape), N > 1: the set cover filter on what the near-duplicate filter keeps
    (-m 5 -l 30 -e 50), ONE grouping with its probes sharded over the GPUs (stage A per shard, stage B with one
    exchange per round).  Every rank builds the same input and runs the same near-duplicate filter first."""
    import torch
    from catch_b200 import probe
    from catch_b200.filter.near_duplicate_filter import NearDuplicateFilterWithMinHash
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    gens = helpers.synthetic_influenza(n_genomes, seed=3)
    groups = [[[seg] for g in gens for seg in g]]
    genomes = helpers.to_genomes(groups)
    cands = helpers.tile_candidates([s for g in groups[0] for s in g], 100, 50)
    T = sum(len(s) for g in groups[0] for s in g)
    np.random.seed(RNG_SEED)
    random.seed(RNG_SEED)
    ndf = NearDuplicateFilterWithMinHash(0.6)
    ndf._ctx = ctx
    kept = ndf.filter([[probe.Probe.from_str(s) for s in cands]], genomes, input_is_grouped=True)
    # the filter returns list(set(...)): an order that depends on the process's string hash seed.  Every rank must
    # hand the SAME list to the sharded filter, so it is put into a process-independent order here
    kept = [sorted(kept[0], key=lambda p: p.seq_str)]
    P = len(kept[0])
    scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=50)
    scf._ctx = ctx
    os.environ['CB_SHARD'] = 'probes'
    best, n_sel = None, 0
    for rep in range(reps + 1):
        np.random.seed(RNG_SEED)
        dist.barrier(device_ids=[local_rank])
        t = time.perf_counter()
        out = scf.filter(kept, genomes, input_is_grouped=True)
        tt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rep > 0 and (best is None or float(tt.item()) < best):
            best = float(tt.item())
        n_sel = len(out[0])
    s = scf.last_stats[0]
    dev = torch.tensor([s['coverage']['ms_total'] + s['setcover']['ms_total']], dtype=torch.float64, device='cuda')
    dist.all_reduce(dev, op=dist.ReduceOp.MAX)
    return {'workload': 'config 3 (influenza shape): %d genomes x 8 segments (T=%d bp), %d probes after the near-duplicate '
                        'filter, -m 5 -l 30 -e 50, probes of the ONE grouping sharded over the GPUs' % (n_genomes, T, P),
            'n_gpus': dist.get_world_size(), 'e2e_ms': best * 1e3, 'device_ms_max_over_ranks': float(dev.item()),
            'pairs_per_s_e2e': P * T / best, 'pairs_per_s_device': P * T / (float(dev.item()) / 1e3), 'selected': n_sel,
            'this_rank': {'scan_ms': s['coverage']['ms_scan_emit'], 'merge_ms': s['coverage']['ms_merge'],
                          'greedy_ms': s['setcover']['ms_greedy'], 'rounds': int(s['setcover']['reserved'][5])}}


def config5_sweep(ctx, kept, genomes, T, cells=None, reps=2):
    """BASELINE config 5: hybridisation sweep on the config-3 input (after the near-duplicate filter)."""
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    os.environ['CB_SHARD'] = 'groups'
    P = len(kept[0])
    out = []
    for m in (0, 2, 5, 10):
        for l in (30, 60, 100):
            if cells is not None and (m, l) not in cells:
                continue
            scf = SetCoverFilter(mismatches=m, lcf_thres=l, cover_extension=50)
            scf._ctx = ctx
            best = None
            for _ in range(reps):
                np.random.seed(RNG_SEED)
                ctx.flush_l2()
                t = time.perf_counter()
                sel = scf.filter(kept, genomes, input_is_grouped=True)
                dt = time.perf_counter() - t
                if best is None or dt < best[0]:
                    best = (dt, dict(scf.last_stats[0]), len(sel[0]))
            dt, s, n_sel = best
            ca, cb = s['coverage'], s['setcover']
            dev_ms = ca['ms_total'] + cb['ms_total']
            out.append({'m': m, 'l': l, 'seeds': '%s k=%d' % (s['seed_mode'], s['k']), 'selected': n_sel,
                        'e2e_ms': round(dt * 1e3, 2), 'device_ms': round(dev_ms, 2),
                        'scan_ms': round(ca['ms_scan_emit'], 2), 'merge_ms': round(ca['ms_merge'], 2),
                        'greedy_ms': round(cb['ms_greedy'], 2), 'rounds': int(cb['reserved'][5]),
                        'pairs_per_s_e2e': P * T / dt, 'pairs_per_s_device': P * T / (dev_ms / 1e3)})
    return out


def config4_vall(ctx, n_taxa, n_genomes, dist=None, local_rank=0, reps=2):
    """BASELINE config 4 shape (V-All): n_taxa independent taxa = groupings, -pl 100 -m 5 -l 30 -e 0, through
    SetCoverFilter.filter(); with a process group the groupings are sharded over the ranks (largest first)."""
    from catch_b200 import probe
    from catch_b200.filter.set_cover_filter import SetCoverFilter
    groups = helpers.synthetic_taxa(n_taxa, n_genomes, seed=4)
    genomes = helpers.to_genomes([[[s] for s in g] for g in groups])
    cands = [list(dict.fromkeys(helpers.tile_candidates(g, 100, 50))) for g in groups]
    probes = [[probe.Probe.from_str(s) for s in c] for c in cands]
    pairs = sum(len(c) * sum(map(len, g)) for c, g in zip(cands, groups))
    scf = SetCoverFilter(mismatches=5, lcf_thres=30, cover_extension=0)
    scf._ctx = ctx
    os.environ['CB_SHARD'] = 'groups'
    best, n_sel = None, 0
    for rep in range(reps + 1):                        # first repetition is the warm-up
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        t = time.perf_counter()
        out = scf.filter(probes, genomes, input_is_grouped=True)
        dt = time.perf_counter() - t
        if dist is not None:
            import torch
            tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if rep > 0 and (best is None or dt < best):
            best = dt
        n_sel = sum(len(o) for o in out)
    dev_ms = sum((s['coverage']['ms_total'] + s['setcover']['ms_total']) for s in scf.last_stats if s and 'coverage' in s)
    world = dist.get_world_size() if dist is not None else 1
    # the same call with the candidates of each grouping as ONE host buffer (catch_b200/probe_batch.py, what
    # the tiling of design.py produces) instead of lists of Probe objects
    from catch_b200.probe_batch import ProbeBatch
    batches = [ProbeBatch(np.frombuffer(''.join(c).encode(), dtype=np.uint8).reshape(len(c), 100)) for c in cands]
    best_b = None
    for rep in range(reps + 1):
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        t = time.perf_counter()
        out_b = scf.filter(batches, genomes, input_is_grouped=True)
        dt = time.perf_counter() - t
        if dist is not None:
            import torch
            tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if rep > 0 and (best_b is None or dt < best_b):
            best_b = dt
    same = [[p.seq_str for p in g] for g in out_b] == [[p.seq_str for p in g] for g in out]
    return {'e2e_batch_ms': best_b * 1e3, 'pairs_per_s_e2e_batch': pairs / best_b, 'batch_output_identical': same,
            'workload': 'config 4 shape (V-All): %d of 300 taxa x %d genomes of 10-30 kb, -pl 100 -m 5 -l 30 -e 0'
                        % (n_taxa, n_genomes),
            'n_gpus': world, 'groupings': n_taxa, 'P_total': sum(map(len, cands)),
            'T_total_bp': sum(sum(map(len, g)) for g in groups), 'pairs': pairs, 'selected': n_sel,
            'e2e_ms': best * 1e3, 'pairs_per_s_e2e': pairs / best, 'this_rank_device_ms': round(dev_ms, 1),
            'sharding': 'single GPU' if world == 1 else
                        'groupings over ranks, largest first; no data-path collective; same taxa at every N (strong)'}


def cluster_f3(ctx, n_taxa=4, n_genomes=333, reps=3, sm_mhz=None):
    """cluster_with_minhash_signatures (the pre-partitioner of design_large.py) on the genomes of a few
    V-All-shape taxa thrown together: sketches (md5 per 12-mer, bottom-100 per genome) and distance rows on the
    device, the search on the host.  The oracle's sketch (hashlib + heapq, one core) is timed on a few genomes."""
    from catch_b200.utils import cluster
    from oracle import oracle
    groups = helpers.synthetic_taxa(n_taxa, n_genomes, seed=4)
    seqs = {'t%d_%d' % (t, i): s for t, g in enumerate(groups) for i, s in enumerate(g)}
    bases = sum(map(len, seqs.values()))
    best, st, clusters = None, None, None
    for _ in range(reps):
        random.seed(RNG_SEED)
        t = time.perf_counter()
        clusters = cluster.cluster_with_minhash_signatures(seqs, threshold=0.15)
        dt = time.perf_counter() - t
        if best is None or dt < best:
            best, st = dt, cluster.cluster_with_minhash_signatures.last_stats
    # md5 of one 12-byte block + the affine map mod 2^31 - 1: 365 SASS instructions per k-mer in
    # sketch_hash_kernel<12> (cuobjdump -sass: 78 IADD3, 75 LEA.HI = rotate-and-add, 74 LOP3, 29 IMAD, ...; 64 md5
    # steps at ~3.6 instructions each), all on the integer pipe.  Ceilings: the rate the same GPU sustains on
    # LOP3 + LEA.HI chains (cb_intop_rate, the scan kernel's ceiling) and the issue limit of 128 lanes per SM-clock
    ops = bases * 365
    hash_s = st['ms_scan_emit'] / 1e3
    peak_ops = ctx.intop_rate()
    issue = 148 * 128 * (sm_mhz or 1965.0) * 1e6
    sample = list(seqs.values())[:8]
    random.seed(RNG_SEED)
    a, b = oracle.sketch_params()
    t = time.perf_counter()
    want = [oracle.sketch(s, 12, 100, a, b) for s in sample]
    cpu_s = time.perf_counter() - t
    got = cluster.SketchFunction(12, 100, a, b).sketch(sample, ctx).signatures().tolist()
    return {'workload': 'genomes of %d V-All-shape taxa x %d (%d sequences, %d bp), k=12 N=100 threshold 0.15, method simple'
                        % (n_taxa, n_genomes, len(seqs), bases),
            'clusters': [len(c) for c in clusters], 'wall_ms': best * 1e3, 'device_ms': st['ms_total'],
            'hash_kernel_ms': st['ms_scan_emit'], 'select_kernel_ms': st['ms_merge'],
            'bases_per_s_e2e': bases / best, 'bases_per_s_hash_kernel': bases / hash_s,
            'int_ops': {'achieved_gops': ops / hash_s / 1e9, 'peak_gops': peak_ops / 1e9, 'frac': ops / hash_s / peak_ops,
                        'issue_limit_gops': issue / 1e9, 'frac_of_issue_limit': ops / hash_s / issue,
                        'instructions_per_kmer': 365},
            'cpu_port': {'bases_per_s': sum(map(len, sample)) / cpu_s, 'cores': 1,
                         'sample': 'first %d genomes (hashlib.md5 + heapq)' % len(sample),
                         'identical': got == [list(x) for x in want]}}


def emit_other_workload(args, ctx):
    """--workload influenza | vall | sweep as the main line (single GPU)."""
    t0 = time.perf_counter()
    line = {'metric': METRIC, 'unit': 'pairs/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic'}
    sampler = ClockSampler(0)
    sampler.start()
    if args.workload == 'vall':
        r = config4_vall(ctx, args.vall_taxa, args.vall_genomes, reps=max(1, args.steps))
        line.update(value=r['pairs_per_s_e2e'], ms_per_step=r['e2e_ms'], config={'workload': r['workload'], 'l2': 'inputs >> L2'},
                    e2e={'value': r['pairs_per_s_e2e'], 'unit': 'pairs/s', 'ms_per_step': r['e2e_ms']}, detail=r)
    else:
        r, kept, genomes, T = config3_influenza(ctx, args.influenza_genomes, reps=max(2, args.steps))
        line.update(value=r['scf']['pairs_per_s_device'], ms_per_step=r['scf']['device_ms'],
                    config={'workload': r['workload'], 'l2': 'flushed between steps (256 MiB memset)'},
                    e2e={'value': r['scf']['pairs_per_s_e2e'], 'unit': 'pairs/s', 'ms_per_step': r['scf']['e2e_ms']},
                    detail=r)
        if args.workload == 'sweep':
            line['sweep'] = config5_sweep(ctx, kept, genomes, T, reps=max(2, args.steps))
    line['clocks'] = sampler.stop()
    line['wall_s'] = round(time.perf_counter() - t0, 1)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='zika', choices=sorted(WORKLOADS) + list(OTHER_WORKLOADS))
    ap.add_argument('--cpu-sample-genomes', type=int, default=60)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip like_for_like and the config 3/4/5 sub-objects')
    ap.add_argument('--shard', default='probes', choices=['groups', 'probes'],
                    help='multi-GPU mode: the probes of ONE grouping split over the GPUs, sharded set cover with one '
                         'exchange per round (strong scaling, default), or one grouping per GPU (weak scaling)')
    ap.add_argument('--influenza-genomes', type=int, default=5000)
    ap.add_argument('--vall-taxa', type=int, default=16)
    ap.add_argument('--vall-genomes', type=int, default=333)
    args = ap.parse_args()
    if args.impl == 'b200' and not os.environ.get('CB_BENCH_PROFILING'):
        args.warmup = max(args.warmup, 3)      # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # plumbing only: barrier, max-over-ranks of the timings (NCCL), small host-side exchanges (gloo)
        dist.init_process_group('cpu:gloo,cuda:nccl')
        # NCCL announces its version on stdout when the communicator is first used; the contract is ONE
        # JSON line on stdout, so the first collective runs with fd 1 pointed at stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from catch_b200 import _lib, parallel, probe
    from catch_b200 import coverage as cov
    from catch_b200.filter.set_cover_filter import SetCoverFilter

    ctx = _lib.Context(local_rank)
    if args.workload in OTHER_WORKLOADS:
        if world > 1:
            raise SystemExit('--workload %s is a single-GPU line; its multi-GPU form is the vall_groups object of the '
                             'default workload' % args.workload)
        emit_other_workload(args, ctx)
        return

    strong = world > 1 and args.shard == 'probes'
    os.environ['CB_SHARD'] = args.shard if world > 1 else 'groups'
    w = make_workload(args.workload, n_groups=1 if (strong or world == 1) else world)
    genomes = helpers.to_genomes([[[s] for s in seqs] for seqs in w['groups_seqs']])
    probes = [[probe.Probe.from_str(s) for s in c] for c in w['groups_cands']]
    scf = SetCoverFilter(**w['scf'])
    scf._ctx = ctx
    # group-sharded run: the grouping this rank owns, by the same assignment SetCoverFilter._filter makes (the
    # deduplicated candidate counts differ per generator seed, so it is a permutation, not g -> rank g)
    my_group = 0
    if world > 1 and not strong:
        sizes = [len(p) * max(1, sum(g.size() for g in tg)) for p, tg in zip(probes, genomes)]
        my_group = parallel.assign_groups(sizes, world).index(rank)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()

    # ---- e2e: the plugin call with host objects (all groupings in, all selections out)
    def e2e_step():
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        ctx.flush_l2()
        t0 = time.perf_counter()
        out = scf.filter(probes, genomes, input_is_grouped=True)
        return time.perf_counter() - t0, out

    # ---- resident: this rank's grouping (all probes uploaded; a rank of a probe-sharded run SCANS only its block)
    my_cands = w['groups_cands'][my_group]
    P = len(my_cands)
    group = cov.PackedGroup(ctx, my_cands, genomes[my_group])
    lo, hi = parallel.shard_bounds(P, world, rank) if strong else (0, P)
    m, lcf, ext = w['scf']['mismatches'], w['scf']['lcf_thres'], w['scf']['cover_extension']

    def resident_step():
        np.random.seed(RNG_SEED)
        ctx.flush_l2()
        plan = cov.SeedPlan(my_cands, m, lcf, 20, may_have_dups=False)
        if strong:
            local, st_a = cov.compute_cover_range(ctx, group, plan, m, lcf, 0, ext, lo, hi)
            parallel.ensure_exchange(ctx, ctx.exchange_required(local), parallel.rng_state_token())
            picks, st_b = ctx.setcover_sharded(local, P, lo, hi)
            local.free()
        else:
            cover, st_a = cov.compute_cover(ctx, group, plan, m, lcf, 0, ext)
            picks, st_b = ctx.setcover(cover, P, None, None)
            cover.free()
        return st_a, st_b, picks

    for _ in range(args.warmup):
        resident_step()
        e2e_step()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    res = []
    for _ in range(args.steps):
        if dist is not None:
            dist.barrier(device_ids=[local_rank])   # ranks of a sharded step start together
        res.append(resident_step())
    barrier()
    e2e = []
    for _ in range(args.steps):
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        e2e.append(e2e_step())
    barrier()
    clocks = sampler.stop()

    ms_res = [a.ms_total + b.ms_total for a, b, _ in res]
    t_res = float(np.mean(ms_res)) / 1e3
    t_e2e = float(np.mean([t for t, _ in e2e]))
    identical_across_ranks = None
    if dist is not None:
        import torch
        t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = t.tolist()
        if strong:
            # every rank must hold the same picks (resident step) and return the same probes (e2e)
            sel = [p.seq_str for p in e2e[-1][1][0]]
            mine = (res[-1][2].tolist(), sel)
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            identical_across_ranks = all(x == everyone[0] for x in everyone)
            if not identical_across_ranks:
                raise SystemExit('ranks disagree on the selection')
    st_a, st_b, picks = res[-1]
    launches = sum(int(a.n_kernel_launches + b.n_kernel_launches) for a, b, _ in res)
    ls = scf.last_stats[my_group]
    launches += args.steps * int(ls['upload_targets']['n_kernel_launches'] + ls['upload_probes']['n_kernel_launches'] +
                                 ls['coverage']['n_kernel_launches'] + ls['setcover']['n_kernel_launches'])

    # ---- roofline of the dominant kernel
    kern = {
        'scan_kernel (K3, count+emit)': (st_a.ms_scan_count + st_a.ms_scan_emit),
        'greedy_rounds_kernel (K6-K8)': st_b.ms_greedy,
        'merge_kernel (K4)': st_a.ms_merge,
        'seed_index (K2)': st_a.ms_seed_index,
        'set-up of stage B (K5: universe, gains, interval index)': st_b.ms_universe,
    }
    dom = max(kern, key=kern.get)
    T = sum(len(s) for s in w['groups_seqs'][my_group])
    E, S = int(st_a.n_intervals), int(st_b.n_picks)
    bits = group.bits
    n_seed_entries = int(st_a.n_seed_entries)
    nw = (w['pl'] + 63) // 64
    hits, surv, lookups = int(st_a.n_candidate_hits), int(st_a.reserved[0]), int(st_a.n_seed_lookups)
    l2_operand_bytes = None
    int_ops = None
    if dom.startswith('scan'):
        # SURVEY.md 8(d), stage A compulsory bytes: target planes (read by the counting pre-pass and by
        # the scan) + one probe record per probe + one 16-byte seed-index entry per distinct seed + 16 B per
        # emitted range.  Everything else the kernel touches (an index entry per candidate hit, a probe
        # record per surviving hit) is re-use served by L1/L2 and is reported as l2_operand_bytes.
        alg_bytes = 2 * T * bits / 8 + (hi - lo) * (bits + 1) * nw * 8 + n_seed_entries * 16 + st_a.n_raw_ranges * 16
        l2_operand_bytes = hits * 16 + surv * (bits * nw * 8 + 8)
        dur = kern[dom] / 1e3
        # integer-op roofline (SURVEY 8d asks for int-op throughput beside HBM): algorithmic 32-bit integer
        # operations of the scan (per looked-up position, per candidate hit, per surviving
        # hit) over the kernel time, against the rate this GPU sustains on independent LOP3/LEA chains, measured now
        kc = (20 + 63) // 64
        ops = lookups * (15 * bits * kc + 8) + hits * (4 + 6 * bits) + \
            surv * (8 * bits * nw + 16 + 22 * (m + 1) + 12)
        peak_ops = ctx.intop_rate()
        int_ops = {'achieved_gops': ops / dur / 1e9, 'peak_gops': peak_ops / 1e9, 'frac': ops / dur / peak_ops,
                   'algorithmic_ops': ops, 'candidate_hits': hits, 'surviving_hits': surv, 'lookups': lookups,
                   'raw_ranges': int(st_a.n_raw_ranges),
                   'peak_source': 'cb_intop_rate measured in this run (integer ALU instructions per second on independent LOP3+LEA.HI chains, all SMs)'}
    elif dom.startswith('greedy'):
        # SURVEY 8(d), stage B: S*P*4 (gain vector per pick) + E*8 (index items touched at least once) + 2*U/8
        alg_bytes = S * P * 4 + E * 8 + 2 * (T / 8)
        dur = kern[dom] / 1e3
    else:
        alg_bytes = st_a.n_raw_ranges * 16 * 3
        dur = kern[dom] / 1e3
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / dur / 1e9 if dur > 0 else 0.0
    traffic, traffic_note = captured_traffic(dom)
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'algorithmic_bytes': alg_bytes, 'peak_source': peak_src,
                'kernel_ms': {k: round(v, 3) for k, v in kern.items()},
                'l2_operand_gbs': (l2_operand_bytes / dur / 1e9) if l2_operand_bytes else None,
                'int_ops': int_ops,
                'traffic_note': traffic_note,
                'note': 'HBM is not what binds this integer path: the scan is bound by the integer ALU pipe (see int_ops '
                        'and the ncu pipe utilisation in profiles/), the greedy loop by grid-barrier and dependent-load '
                        'latency per round; frac is reported against HBM as the contract asks'}

    if strong:
        par = ('probes of ONE grouping split over the GPUs: stage A per shard, stage B with sharded gains / interval '
               'index, replicated universe and one exchange per round through peer-mapped memory (csrc/rounds.cu); '
               'torch.distributed carries only control data')
    elif world > 1:
        par = 'one grouping per GPU (independent set-cover instances), no data-path collective'
    else:
        par = 'single GPU'
    out = {
        'metric': METRIC,
        'value': w['pairs'] / t_res, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': t_res * 1e3, 'higher_is_better': True,
        'scaling': 'strong' if (strong or world == 1) else 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': w['desc'] + ('' if (world == 1 or strong) else ' x %d independent groupings' % world),
                   'P_per_group': P, 'T_bp_per_group': T, 'pairs': w['pairs'], 'intervals': E,
                   'picks': S, 'l2': 'flushed between steps (256 MiB memset)',
                   'parallelism': par},
        'e2e': {'value': w['pairs'] / t_e2e, 'unit': 'pairs/s', 'ms_per_step': t_e2e * 1e3,
                'h2d_bytes_per_step': int(ls['h2d_bytes']) * world,
                'd2h_bytes_per_step': int(ls['d2h_bytes']) * world},
        'gpu_launches': launches * world,
        'clocks': clocks,
        'roofline': roofline,
        'stages_ms': {'coverage': st_a.as_dict(), 'setcover': st_b.as_dict(),
                      'e2e_host': {k: round(v, 2) for k, v in ls.get('host_ms', {}).items()},
                      'e2e_group_wall_ms': round(ls.get('wall_s', 0) * 1e3, 2)},
    }
    if strong:
        out['multi_gpu'] = {'exchange_ranks_seen': int(getattr(ctx, 'exchange_n_ranks', 0)),
                            'selection_identical_on_every_rank': identical_across_ranks,
                            'greedy_rounds': int(st_b.reserved[5]), 'list_rebuilds': int(st_b.reserved[4]),
                            'this_rank_probes': [int(lo), int(hi)], 'this_rank_intervals': E}
    group.free()

    # ---- extras: same-input comparison with the CPU leg, the other BASELINE configs
    extras_ok = args.workload == 'zika' and not args.no_extras
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu, lfl = cpu_and_like_for_like(args, ctx, args.workload)
            out['cpu_baseline'] = cpu
            if not args.no_extras:
                out['like_for_like'] = lfl
        except Exception as e:                      # the main line must survive a failing extra
            out['cpu_baseline_error'] = repr(e)
    if extras_ok and world == 1:
        configs = {}
        try:
            r3, kept, g3, T3 = config3_influenza(ctx, args.influenza_genomes)
            configs['config3_influenza'] = r3
            configs['config5_sweep'] = config5_sweep(ctx, kept, g3, T3, cells={(0, 100), (2, 60), (5, 30), (10, 30)})
            del kept, g3
        except Exception as e:
            configs['config3_error'] = repr(e)
        try:
            configs['config4_vall'] = config4_vall(ctx, args.vall_taxa, args.vall_genomes)
        except Exception as e:
            configs['config4_error'] = repr(e)
        try:
            configs['cluster_f3'] = cluster_f3(ctx, sm_mhz=(clocks or {}).get('sm_mhz'))
        except Exception as e:
            configs['cluster_f3_error'] = repr(e)
        out['configs'] = configs
    if extras_ok and world > 1:
        try:
            out['vall_groups'] = config4_vall(ctx, args.vall_taxa, args.vall_genomes, dist=dist, local_rank=local_rank)
        except Exception as e:
            out['vall_groups_error'] = repr(e)
        try:
            out['influenza_shard'] = config3_sharded(ctx, args.influenza_genomes, dist, local_rank)
        except Exception as e:
            out['influenza_shard_error'] = repr(e)
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
