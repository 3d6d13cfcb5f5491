#!/usr/bin/env python
"""Benchmark of the SetCoverFilter hot path (BASELINE.json metric: candidate-probe x target-bp / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload zika] [--impl reference]

A step is one pass of the hot path (stage A coverage + stage B greedy set cover) over one batch
of synthetic genomes.  Two numbers per run:
  value : inputs already packed and resident in HBM when the timed region starts (device
          timeline, CUDA events on the library's stream);
  e2e   : the same metric through the plugin call a user makes, SetCoverFilter.filter(), with
          host Probe/Genome objects; host packing, host->device copies, both stages and the
          device->host read of the selection are all inside the timed region.
The CPU oracle (oracle/) is executed only for `cpu_baseline` and `--impl reference`.
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

WORKLOADS = {
    # name: (n_genomes, genome_len, divergence, generator seed, probe_length, probe_stride, filter kwargs)
    'plumbing': dict(n_genomes=20, length=5000, div=0.03, seed=1, pl=75, ps=50,
                     scf=dict(mismatches=0, lcf_thres=75, cover_extension=0),
                     desc='config 1: 20 x 5 kb, -pl 75 -m 0 -e 0'),
    'zika': dict(n_genomes=500, length=11000, div=0.03, seed=2, pl=75, ps=50,
                 scf=dict(mismatches=2, lcf_thres=60, cover_extension=50),
                 desc='config 2 (Zika-scale): 500 x 11 kb, -pl 75 -m 2 -l 60 -e 50'),
}


def make_workload(name, n_genomes=None, n_groups=1):
    """n_groups independent groupings of the named shape (generator seeds seed, seed+1, ...):
    one grouping per GPU in the multi-GPU runs (weak scaling over independent set-cover instances,
    the V-All structure of many taxa)."""
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


# ------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU restatement of the reference path (oracle port -- the reference
    itself is Python and does not travel to the GPU box) on the host cores, bounded sample."""
    if rank != 0:
        return
    from oracle import oracle as O
    w = make_workload(args.workload, n_genomes=args.cpu_sample_genomes)
    threads = O.num_threads()
    groups = [[[s] for s in w['seqs']]]

    def step():
        np.random.seed(RNG_SEED)
        random.seed(RNG_SEED)
        t0 = time.perf_counter()
        O.set_cover_filter([w['cands']], groups, w['scf']['mismatches'], w['scf']['lcf_thres'], 0, 1.0,
                           w['scf']['cover_extension'], 20, n_threads=threads)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    t = sum(times) / len(times)
    v = w['pairs'] / t
    sample = 'first %d genomes of %s (P=%d, T=%d bp)' % (w['n_genomes'], args.workload, len(w['cands']),
                                                          sum(map(len, w['seqs'])))
    print(json.dumps({
        'impl': 'reference', 'metric': 'candidate-probe x target-bp / s through SetCoverFilter',
        'value': v, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload]['desc'], 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def cpu_baseline(args):
    from oracle import oracle as O
    w = make_workload(args.workload, n_genomes=args.cpu_sample_genomes)
    threads = O.num_threads()
    np.random.seed(RNG_SEED)
    random.seed(RNG_SEED)
    t0 = time.perf_counter()
    O.set_cover_filter([w['cands']], [[[s] for s in w['seqs']]], w['scf']['mismatches'],
                       w['scf']['lcf_thres'], 0, 1.0, w['scf']['cover_extension'], 20, n_threads=threads)
    t = time.perf_counter() - t0
    return {'value': w['pairs'] / t, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
            'sample': 'first %d genomes of %s (P=%d, T=%d bp), %.1f s; stage A on %d threads, stage B on 1 '
                      '(as the reference: one process per grouping)' % (
                          w['n_genomes'], args.workload, len(w['cands']), sum(map(len, w['seqs'])), t, threads)}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='zika', choices=sorted(WORKLOADS))
    ap.add_argument('--cpu-sample-genomes', type=int, default=60)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--shard', default='groups', choices=['groups', 'probes'],
                    help='multi-GPU mode: one grouping per GPU (weak scaling, default) or the probes of ONE '
                         'grouping split over the GPUs with an NCCL all-gather of the coverage (strong scaling)')
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
        # plumbing only: barrier, max-over-ranks of the timings (NCCL), exchange of the selected
        # ids between ranks (gloo, a few KB of Python objects)
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

    from catch_b200 import _lib, probe
    from catch_b200 import coverage as cov
    from catch_b200.filter.set_cover_filter import SetCoverFilter

    strong = world > 1 and args.shard == 'probes'
    os.environ['CB_SHARD'] = args.shard
    w = make_workload(args.workload, n_groups=1 if strong else world)
    ctx = _lib.Context(local_rank)
    genomes = helpers.to_genomes([[[s] for s in seqs] for seqs in w['groups_seqs']])
    probes = [[probe.Probe.from_str(s) for s in c] for c in w['groups_cands']]
    scf = SetCoverFilter(**w['scf'])
    scf._ctx = ctx
    # group-sharded run: the grouping this rank owns, by the same assignment SetCoverFilter._filter makes (the
    # deduplicated candidate counts differ per generator seed, so it is a permutation, not g -> rank g)
    my_group = 0
    if world > 1 and not strong:
        from catch_b200 import parallel as _par
        sizes = [len(p) * max(1, sum(g.size() for g in tg)) for p, tg in zip(probes, genomes)]
        my_group = _par.assign_groups(sizes, world).index(rank)

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

    # ---- resident: this rank's grouping (or its block of the probes) packed in HBM before the timed region
    my_cands = w['groups_cands'][my_group]
    if strong:
        from catch_b200 import parallel
        parallel.ensure_comm(ctx)
        lo, hi = parallel.shard_bounds(len(my_cands), world, rank)
        group = cov.PackedGroup(ctx, my_cands[lo:hi], genomes[my_group])
    else:
        lo, hi = 0, len(my_cands)
        group = cov.PackedGroup(ctx, my_cands, genomes[my_group])

    def resident_step():
        np.random.seed(RNG_SEED)
        ctx.flush_l2()
        plan = cov.SeedPlan(my_cands, w['scf']['mismatches'], w['scf']['lcf_thres'], 20)
        if strong:
            t0 = time.perf_counter()
            so = np.ascontiguousarray(plan.seed_off[lo:hi + 1] - plan.seed_off[lo])
            sp = np.ascontiguousarray(plan.seed_pos[plan.seed_off[lo]:max(plan.seed_off[hi], plan.seed_off[lo] + 1)])
            local, st_a = ctx.coverage(group.probes, group.targets, w['scf']['mismatches'], w['scf']['lcf_thres'], 0,
                                       w['scf']['cover_extension'], plan.k, so, sp)
            tg = time.perf_counter()
            cover = ctx.cover_allgather(local, lo, len(my_cands))
            st_a.ms_total += (time.perf_counter() - tg) * 1e3        # the exchange is part of the step
            local.free()
        else:
            cover, st_a = cov.compute_cover(ctx, group, plan, w['scf']['mismatches'], w['scf']['lcf_thres'], 0,
                                            w['scf']['cover_extension'])
        picks, st_b = ctx.setcover(cover, len(my_cands), None, None)
        cover.free()
        return st_a, st_b, picks

    for _ in range(args.warmup):
        resident_step()
        e2e_step()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    res = [resident_step() for _ in range(args.steps)]
    barrier()
    e2e = [e2e_step() for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop()

    ms_res = [a.ms_total + b.ms_total for a, b, _ in res]
    t_res = float(np.mean(ms_res)) / 1e3
    t_e2e = float(np.mean([t for t, _ in e2e]))
    if dist is not None:
        import torch
        t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = t.tolist()
    st_a, st_b, picks = res[-1]
    launches = sum(int(a.n_kernel_launches + b.n_kernel_launches) for a, b, _ in res)
    ls = scf.last_stats[my_group]
    launches += args.steps * int(ls['upload_targets']['n_kernel_launches'] + ls['upload_probes']['n_kernel_launches'] +
                                 ls['coverage']['n_kernel_launches'] + ls['setcover']['n_kernel_launches'])

    # ---- roofline of the dominant kernel
    kern = {
        'scan_kernel (K3, count+emit)': (st_a.ms_scan_count + st_a.ms_scan_emit),
        'greedy_kernel (K6-K8)': st_b.ms_greedy,
        'merge_kernel (K4)': st_a.ms_merge,
        'seed_index (K2)': st_a.ms_seed_index,
    }
    dom = max(kern, key=kern.get)
    P, T = len(my_cands), sum(len(s) for s in w['groups_seqs'][my_group])
    E, S = int(st_a.n_intervals), int(st_b.n_picks)
    bits = group.bits
    n_seed_entries = int(st_a.n_seed_entries)
    l2_operand_bytes = None
    if dom.startswith('scan'):
        # SURVEY.md 8(d), stage A compulsory bytes: target planes (read by the counting pre-pass and by
        # the scan) + one probe record per probe + one seed-index entry per distinct seed + 16 B per
        # emitted range.  Everything else the kernel touches (index entries and probe records per
        # candidate hit) is re-use served by L1/L2 and is reported separately as l2_operand_bytes.
        nw = (w['pl'] + 63) // 64
        alg_bytes = 2 * T * bits / 8 + P * (bits + 1) * nw * 8 + n_seed_entries * 8 + st_a.n_raw_ranges * 16
        l2_operand_bytes = st_a.n_candidate_hits * (8 + (bits + 1) * nw * 8)
        dur = kern[dom] / 1e3
    elif dom.startswith('greedy'):
        # SURVEY 8(d), stage B: S*P*4 (gain vector per pick) + E*16 (index items touched at least once) + 2*U/8
        alg_bytes = S * P * 4 + E * 16 + 2 * (T / 8)
        dur = kern[dom] / 1e3
    else:
        alg_bytes = st_a.n_raw_ranges * 16 * 3
        dur = kern[dom] / 1e3
    # measured DRAM traffic of the same kernel on the same workload (one `ncu --set full` capture,
    # profiles/traffic_r01b.json); None for other workloads
    traffic, ncu_facts = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic_r01b.json')) as f:
            tj = json.load(f)
        if tj.get('workload') == args.workload:
            key = 'scan_kernel' if dom.startswith('scan') else 'greedy_kernel'
            traffic = tj['dram_bytes_per_launch'].get(key)
            ncu_facts = tj.get('ncu', {}).get(key)
    except Exception:
        pass
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / dur / 1e9 if dur > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'algorithmic_bytes': alg_bytes, 'peak_source': peak_src,
                'kernel_ms': {k: round(v, 3) for k, v in kern.items()},
                'l2_operand_gbs': (l2_operand_bytes / dur / 1e9) if l2_operand_bytes else None,
                'ncu': ncu_facts,
                'note': 'HBM is not what binds this integer path: the scan is bound by the integer ALU pipe '
                        '(ncu: ALU pipe 70 %, issue slots 63 %, operands L2-resident per grouping), the greedy loop '
                        'by grid-barrier and dependent-load latency; frac is reported against HBM as the contract '
                        'asks, see DESIGN.md section 5 and profiles/README_r01.md'}

    out = {
        'metric': 'candidate-probe x target-bp / s through SetCoverFilter',
        'value': w['pairs'] / t_res, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': t_res * 1e3, 'higher_is_better': True,
        'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': w['desc'] + ('' if (world == 1 or strong) else ' x %d independent groupings' % world),
                   'P_per_group': P, 'T_bp_per_group': T, 'pairs': w['pairs'], 'intervals': E,
                   'picks': S, 'l2': 'flushed between steps (256 MiB memset)',
                   'parallelism': 'single GPU' if world == 1 else
                   ('probes of one grouping split over the GPUs, NCCL all-gather of the coverage, greedy replicated'
                    if strong else
                    'one grouping per GPU (independent set-cover instances), no data-path collective')},
        'e2e': {'value': w['pairs'] / t_e2e, 'unit': 'pairs/s', 'ms_per_step': t_e2e * 1e3,
                'h2d_bytes_per_step': int(scf.last_stats[my_group]['h2d_bytes']) * world,
                'd2h_bytes_per_step': int(scf.last_stats[my_group]['d2h_bytes']) * world},
        'gpu_launches': launches * world,
        'clocks': clocks,
        'roofline': roofline,
        'stages_ms': {'coverage': st_a.as_dict(), 'setcover': st_b.as_dict(),
                      'e2e_host': {k: round(v, 2) for k, v in scf.last_stats[my_group].get('host_ms', {}).items()},
                      'e2e_group_wall_ms': round(scf.last_stats[my_group].get('wall_s', 0) * 1e3, 2)},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(args)
        print(json.dumps(out))
    group.free()
    if dist is not None:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
