"""SetCoverFilter on the device: the drop-in for catch/filter/set_cover_filter.py:195-930.

Same constructor, attributes and `_filter(input, target_genomes_grouped)` contract as the
reference (:199-213, :902-930): for every grouping it computes which bases of the grouping's
target genomes each candidate probe covers (stage A, reference `_make_sets` :359-470), then picks
probes by greedy multi-universe set cover (stage B, `set_cover.approx_multiuniverse`), and
returns the SAME Probe objects it was given, in the order the reference would emit them.

Host work kept here on purpose: the seed choice (it replays numpy's global RNG exactly as the
reference consumes it), duplicate handling, ranks/universe_p bookkeeping and the Python-set
ordering of the output.  Everything else is in libcatchb200.so; there is no CPU fallback.
"""
import logging
import os
import pickle
import threading
import time

import numpy as np

from catch_b200 import _lib
from catch_b200 import coverage as cov
from catch_b200 import parallel
from catch_b200.filter.base_filter import BaseFilter
from catch_b200.probe_batch import ProbeBatch
from catch_b200.utils import seq_io

logger = logging.getLogger(__name__)


def set_max_num_processes_for_set_cover_instances(max_num_processes=8):
    """Interface compatibility (set_cover_filter.py:66-80); groupings are solved on the GPU one
    after the other, no process pool exists."""
    global _sc_max_num_processes
    _sc_max_num_processes = max_num_processes


set_max_num_processes_for_set_cover_instances()

_RC = str.maketrans('ACGT', 'TGCA')


def _reverse_complement(s):
    # rc_map.get(b, b): anything outside ACGT maps to itself (set_cover_filter.py:514-520)
    return s[::-1].translate(_RC)


class _LazyStrs:
    """The seq_str of every probe of a list, materialised on first use: the common path (no
    duplicates, no ranks) never needs the Python strings, only the gathered bytes."""

    def __init__(self, probes):
        self._probes = probes
        self._strs = None

    def _get(self):
        if self._strs is None:
            self._strs = self._probes.strs() if isinstance(self._probes, ProbeBatch) else \
                [p.seq_str for p in self._probes]
        return self._strs

    def __len__(self):
        return len(self._probes)

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, i):
        return self._get()[i]


class _DrawChain:
    """Seed draws of a group-sharded run.  numpy's global RNG state is a token that travels along the
    groupings in order: the owner of grouping g receives it from the owner of g-1, makes the draws of
    g (first for _make_sets, then the tolerant ones for _make_ranks, set_cover_filter.py:824-827) and
    passes it on.  That chain is inherently sequential, so it runs on its own thread, decoupled from
    the uploads and device work of the rank's groupings: a rank that is busy on the GPU must not
    hold up the owners of the groupings that follow.  Only this thread touches np.random while the
    chain is running."""

    def __init__(self, flt, input, owner, rank):
        self.flt, self.owner, self.rank, self.n_groups = flt, owner, rank, len(input)
        self.mine = [g for g in range(len(input)) if owner[g] == rank]
        # lengths first, on the calling thread and on all ranks at once, so that a hop of the chain
        # is nothing but the draw itself
        self.lengths = {g: cov.probe_lengths(input[g] if isinstance(input[g], (list, tuple, ProbeBatch)) else list(input[g]))
                        for g in self.mine}
        self.events = {g: threading.Event() for g in self.mine}
        self.results, self.sends, self.error = {}, [], None
        self.thread = threading.Thread(target=self._run, name='cb-draw-chain', daemon=True)
        self.thread.start()

    def _run(self):
        flt = self.flt
        for g in self.mine:
            try:
                if g > 0 and self.owner[g - 1] != self.rank:
                    parallel.rng_recv(self.owner[g - 1])
                drawn = drawn_tol = None
                lens = self.lengths[g]
                try:
                    if len(lens):
                        drawn = cov.draw_seeds(lens, flt.mismatches, flt.lcf_thres, flt.kmer_probe_map_k)
                        if flt._needs_ranks():
                            drawn_tol = cov.draw_seeds(lens, flt.mismatches_tolerant, flt.lcf_thres_tolerant,
                                                       flt.kmer_probe_map_k)
                finally:
                    # the token moves on even if this grouping's parameters are rejected, so that the
                    # other ranks are not left waiting
                    if g + 1 < self.n_groups and self.owner[g + 1] != self.rank:
                        self.sends.append(parallel.rng_isend(self.owner[g + 1]))
                self.results[g] = (drawn, drawn_tol)
            except BaseException as e:
                self.error = e
            self.events[g].set()

    def get(self, g, true_lengths=None):
        self.events[g].wait()
        if self.error is not None:
            raise self.error
        return self.results.pop(g)

    def finish(self):
        self.thread.join()
        for req, _buf in self.sends:
            req.wait()
        if self.n_groups:
            parallel.rng_broadcast(self.owner[-1])            # everyone ends where a single process would


class _LengthsNotUniform(RuntimeError):
    """The length samples of a draw replay missed probes of another length (see _DrawReplay)."""


class _DrawReplay:
    """Seed draws of a group-sharded run without any communication: EVERY rank makes the draws of ALL groupings,
    in grouping order, on a helper thread, and keeps those of the groupings it owns.  numpy's stream is a chain
    (grouping g continues where g-1 stopped, set_cover_filter.py:824-827), and the number of raw outputs a grouping
    consumes is only known by generating them (rejection sampling) -- but generating them is cheap now (AVX-512
    replay, ~1.6 ms for 128 k probes x 20 draws), cheaper than a hop of the token chain above (draw + send + wake-up),
    and the draws of grouping g are ready on every rank g x 1.6 ms after the start instead of g hops later.  Every
    rank ends with the state a single process would have: no broadcast.
    Only the lengths of the probes enter the draws.  A rank does not look at the probes of groupings it does not own;
    it takes them to be as long as a few evenly spaced samples of the list say (candidate probes are: one length),
    and measures the whole list only when the samples disagree.  The owner checks the assumption against the lengths
    it gathers and fails loudly if it does not hold (CB_DRAW=chain selects the token chain, which assumes nothing)."""

    def __init__(self, flt, input, owner, rank):
        self.flt, self.owner, self.rank, self.n_groups = flt, owner, rank, len(input)
        self.mine = set(g for g in range(len(input)) if owner[g] == rank)
        self.lengths = [self._lengths_of(p) for p in input]
        self.events = {g: threading.Event() for g in self.mine}
        self.results, self.error = {}, {}
        self.thread = threading.Thread(target=self._run, name='cb-draw-replay', daemon=True)
        self.thread.start()

    @staticmethod
    def _lengths_of(probes):
        """(n, L) when all probes are taken to be L long, else the int32 array of all lengths."""
        if isinstance(probes, ProbeBatch):
            return len(probes), probes.probe_length
        if not isinstance(probes, (list, tuple)):
            probes = list(probes)
        n = len(probes)
        if n == 0:
            return 0, 0
        idx = sorted(set(int(round(x)) for x in np.linspace(0, n - 1, num=min(n, 48)).tolist()))
        sample = cov.probe_lengths([probes[i] for i in idx])
        if bool((sample == sample[0]).all()):
            return n, int(sample[0])
        return cov.probe_lengths(probes)

    def assumed_lengths(self, g):
        lens = self.lengths[g]
        return np.full(lens[0], lens[1], dtype=np.int32) if isinstance(lens, tuple) else lens

    def _run(self):
        flt = self.flt
        for g in range(self.n_groups):
            drawn = drawn_tol = None
            try:
                lens = self.assumed_lengths(g)
                if len(lens):
                    drawn = cov.draw_seeds(lens, flt.mismatches, flt.lcf_thres, flt.kmer_probe_map_k)
                    if flt._needs_ranks():
                        drawn_tol = cov.draw_seeds(lens, flt.mismatches_tolerant, flt.lcf_thres_tolerant,
                                                   flt.kmer_probe_map_k)
            except BaseException as e:             # parameters rejected: raised where the grouping is processed
                self.error[g] = e
            if g in self.mine:
                self.results[g] = (drawn, drawn_tol)
                self.events[g].set()

    def get(self, g, true_lengths=None):
        self.events[g].wait()
        if g in self.error:
            raise self.error[g]
        if true_lengths is not None and not np.array_equal(np.asarray(true_lengths, dtype=np.int64),
                                                           self.assumed_lengths(g).astype(np.int64)):
            raise _LengthsNotUniform("grouping %d: probes of different lengths were not seen by the length samples "
                                     "of the draw replay; run with CB_DRAW=chain" % g)
        return self.results.pop(g)

    def finish(self):
        self.thread.join()


class _Prefetch:
    """Host work of the NEXT grouping while the current one is on the device.  A helper thread gathers the
    probe sequences and the target sequences of the groupings this process owns, in order, into a second
    pair of page-locked staging buffers (cb_host_buffer slots 2 + 2k and 3 + 2k of pair k = 0, 1; slots 0 and 1
    stay with the calls the loop itself makes, e.g. the scans behind the ranks); the filter loop picks
    each grouping up when it gets there.  The device calls of the loop release the GIL (ctypes), so the
    gather of grouping g+1 overlaps the scan and set cover of g.  A pair is reused only after the upload
    that reads it has returned (release()).  The staging buffers are sized up front on the calling thread,
    so the helper never calls into the library.
    ON by default (CB_PREFETCH=0 turns it off).  Measured on the V-All shape (16 taxa, one B200): lists of Probe
    objects 262 -> 234 ms, ProbeBatch input 198 -> 176 ms.  The gather of a list of Probe objects needs the GIL
    for its attribute pass; it hands the lock over every 1024 objects (csrc/fastpack.c), so the main thread is not
    kept from issuing its next device call.  The larger remedy for that host cost is not to have the objects at
    all: ProbeBatch."""

    def __init__(self, ctx, groups):
        """groups: list of (group index, probe list, genomes) in processing order."""
        import queue
        self.ctx, self.groups = ctx, groups
        p_max = t_max = 0
        for _g, probes, genomes in groups:
            if len(probes):
                p_max = max(p_max, int(cov.probe_lengths(probes).sum(dtype=np.int64)))
            t_max = max(t_max, sum(g.size() for g in genomes))
        for pair in (0, 1):
            ctx.host_buffer(2 + 2 * pair, p_max + 64)
            ctx.host_buffer(3 + 2 * pair, t_max + 64)
        self.free = [threading.Semaphore(1), threading.Semaphore(1)]
        self.q = queue.Queue()
        self.stop = False
        self.thread = threading.Thread(target=self._run, name='cb-prefetch', daemon=True)
        self.thread.start()

    def _run(self):
        for k, (g, probes, genomes) in enumerate(self.groups):
            self.free[k % 2].acquire()
            if self.stop:
                return
            try:
                gathered = cov.gather_staged(self.ctx, 2 + 2 * (k % 2), probes) if len(probes) else None
                staged = cov.stage_targets(self.ctx, 3 + 2 * (k % 2), genomes) if len(probes) else None
                self.q.put((g, gathered, staged, None))
            except BaseException as e:          # noqa: BLE001 -- re-raised by the consumer
                self.q.put((g, None, None, e))

    def get(self, g):
        got, gathered, staged, err = self.q.get()
        assert got == g
        if err is not None:
            raise err
        return gathered, staged

    def release(self, k):
        self.free[k % 2].release()

    def close(self):
        self.stop = True
        for s in self.free:
            s.release()
        self.thread.join()


class SetCoverFilter(BaseFilter):
    def __init__(self, mismatches, lcf_thres, island_of_exact_match=0, mismatches_tolerant=None,
                 lcf_thres_tolerant=None, island_of_exact_match_tolerant=None,
                 custom_cover_range_fn=None, custom_cover_range_tolerant_fn=None, identify=False,
                 avoided_genomes=[], coverage=1.0, cover_extension=0, kmer_probe_map_k=20,
                 kmer_probe_map_use_native_dict=False):
        if custom_cover_range_fn is not None or custom_cover_range_tolerant_fn is not None:
            # the reference loads an arbitrary Python predicate (set_cover_filter.py:296-299);
            # that cannot run inside a CUDA kernel
            raise NotImplementedError("custom hybridization functions are Python callables and "
                                      "are not supported by the device implementation")
        self.mismatches = mismatches
        self.lcf_thres = lcf_thres
        self.island_of_exact_match = island_of_exact_match
        # reference keeps callables here; the device path carries the parameters instead
        self.cover_range_fn = ('lcf', mismatches, lcf_thres, island_of_exact_match)
        if not mismatches_tolerant:
            mismatches_tolerant = mismatches
        if not lcf_thres_tolerant:
            lcf_thres_tolerant = lcf_thres
        if not island_of_exact_match_tolerant:
            island_of_exact_match_tolerant = island_of_exact_match
        self.mismatches_tolerant = mismatches_tolerant
        self.lcf_thres_tolerant = lcf_thres_tolerant
        self.island_of_exact_match_tolerant = island_of_exact_match_tolerant
        self.cover_range_tolerant_fn = ('lcf', mismatches_tolerant, lcf_thres_tolerant,
                                        island_of_exact_match_tolerant)
        if identify:
            if (coverage <= 1.0 and coverage >= 0.25) or (coverage > 1 and coverage >= 5000):
                logger.warning("Identification is enabled but the required coverage is high; "
                               "generally coverage should be small when performing identification")
        self.identify = identify
        self.avoided_genomes = avoided_genomes
        self.coverage = coverage
        self.cover_extension = cover_extension
        self.kmer_probe_map_k = kmer_probe_map_k
        self.kmer_probe_map_use_native_dict = kmer_probe_map_use_native_dict
        self.requires_probe_groupings = True
        self._force_num_processes = None       # accepted and ignored (tests set it)
        self._ctx = None
        self._tls = threading.local()          # per worker thread: the context it drives, its host timings
        self.last_stats = []                   # one dict per grouping of the last _filter call

    # ------------------------------------------------------------------ helpers
    def __getstate__(self):
        # a filter travels without its device context and per-thread state (both are re-made on first use)
        d = dict(self.__dict__)
        d.pop('_tls', None)
        d['_ctx'] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._tls = threading.local()

    def _context(self):
        ctx = getattr(self._tls, 'ctx', None)
        if ctx is not None:
            return ctx
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _make_universe_p(self, target_genomes):
        """set_cover_filter.py:761-792."""
        if self.coverage <= 1.0:
            return np.full(len(target_genomes), float(self.coverage), dtype=np.float64)
        p = np.empty(len(target_genomes), dtype=np.float64)
        for j, g in enumerate(target_genomes):
            size = g.size()
            p[j] = float(min(self.coverage, size)) / size
        return p

    def _tolerant_bp_covered(self, ctx, probe_strs, plan, sequences):
        """Sum over `sequences` (any iterable, consumed as a stream) and their reverse complements of the bp
        each probe covers under the tolerant parameters (set_cover_filter.py:472-529): ranges merged per
        sequence, no cover extension.  Each sequence and each reverse complement is packed as its own
        'genome' so merging stays per sequence.  The sequences go through the device in batches of bounded
        size (a host genome given to --avoid-genomes is gigabases long; the reference, too, looks at one
        sequence at a time, :700-707), so neither the 2^32-bit universe limit nor memory depends on the size
        of the FASTA."""
        bp = np.zeros(len(probe_strs), dtype=np.int64)
        for batch in cov.sequence_batches(sequences, max_bases=1 << 29):
            seqs = []
            for s in batch:
                seqs.append([s])
                seqs.append([_reverse_complement(s)])
            group = cov.PackedGroup(ctx, probe_strs, seqs)
            try:
                cover, st = cov.compute_cover(ctx, group, plan, self.mismatches_tolerant,
                                              self.lcf_thres_tolerant, self.island_of_exact_match_tolerant, 0)
                pid, _, start, end = ctx.cover_export(cover)
                cover.free()
            finally:
                group.free()
            if len(pid):
                np.add.at(bp, pid, end - start)
        return bp

    def _needs_ranks(self):
        return bool(self.identify) or len(self.avoided_genomes) > 0

    def _make_ranks(self, ctx, probe_strs, plan, target_genomes_grouped):
        """set_cover_filter.py:614-735.  `plan` is the tolerant seed plan: the reference builds ONE
        k-mer map per call (:672-681) and scans every grouping / avoided genome with it."""
        n = len(probe_strs)
        first = np.zeros(n, dtype=np.int64)
        second = np.zeros(n, dtype=np.int64)
        if self.identify:
            hits = np.zeros(n, dtype=np.int64)
            for genomes in target_genomes_grouped:
                bp = self._tolerant_bp_covered(ctx, probe_strs, plan, (s for g in genomes for s in g.seqs))
                hits += (bp >= 1)
            second = hits
        avoided_bp = np.zeros(n, dtype=np.int64)
        for path in self.avoided_genomes:
            avoided_bp += self._tolerant_bp_covered(ctx, probe_strs, plan, seq_io.iterate_fasta(path))
        if plan.rep is not None:
            # dicts keyed by sequence: all duplicates share the values of their representative
            rep = np.asarray(plan.rep)
            second = second[rep]
            avoided_bp = avoided_bp[rep]
        mask = avoided_bp > 0
        first[mask] = 1
        second = np.where(mask, avoided_bp, second)
        tuples = sorted(set(zip(first.tolist(), second.tolist())))
        idx = {t: i for i, t in enumerate(tuples)}
        return np.array([idx[(a, b)] for a, b in zip(first.tolist(), second.tolist())], dtype=np.int32)

    # ------------------------------------------------------------------ the filter
    def _filter(self, input, target_genomes_grouped):
        self.last_stats = [None] * len(input)
        # multi-GPU: groupings are independent instances and are sharded over the ranks; when there
        # are fewer groupings than ranks, the PROBES of each grouping are sharded instead and the
        # per-probe coverage is all-gathered (parallel.ensure_comm / cb_cover_allgather)
        rank, world_size, _ = parallel.world()
        sharded = parallel.active()
        if sharded and self._shard_probes(len(input), world_size):
            return self._filter_probe_sharded(input, target_genomes_grouped, rank, world_size)
        if sharded:
            sizes = [len(p) * max(1, sum(g.size() for g in tg)) for p, tg in zip(input, target_genomes_grouped)]
            owner = parallel.assign_groups(sizes, world_size)
        else:
            owner = [rank] * len(input)
        local = {}
        chain = None
        if sharded:
            chain = (_DrawChain if os.environ.get('CB_DRAW', 'replay') == 'chain' else _DrawReplay)(self, input, owner, rank)
        failure = None
        mine = [g for g in range(len(input)) if owner[g] == rank]
        # Up to three groupings at a time (CB_PIPELINE=1: one after the other): see _run_pipelined.  V-All shape, 16
        # taxa, one B200: 1 / 2 / 3 workers = 219 / 212 / 188 ms from lists of Probe objects, 173 / 153 / 147 ms from
        # ProbeBatch input
        n_workers = min(len(mine), max(1, int(os.environ.get('CB_PIPELINE', '3'))))
        if n_workers >= 2 and isinstance(self._context(), _lib.Context):
            state0 = np.random.get_state()
            try:
                local, failure = self._run_pipelined(input, target_genomes_grouped, mine, chain, n_workers)
                mine = []                            # done
            except _LengthsNotUniform:
                if sharded:
                    raise
                np.random.set_state(state0)          # rare (mixed probe lengths): the plain loop measures every list
                local, failure = {}, None
            if failure is not None and not sharded:
                raise failure
        prefetch = None
        if len(mine) >= 2 and cov._fastpack is not None and hasattr(self._context(), 'host_buffer') and \
                os.environ.get('CB_PREFETCH', '1') == '1':
            as_lists = {g: (input[g] if isinstance(input[g], (list, tuple, ProbeBatch)) else list(input[g])) for g in mine}
            prefetch = _Prefetch(self._context(), [(g, as_lists[g], target_genomes_grouped[g]) for g in mine])
        try:
            for k, group_i in enumerate(mine):
                # The seed draws consume numpy's global RNG per grouping, in grouping order
                # (set_cover_filter.py:824-827).  A sharded run gets them from the draw chain (every
                # grouping sees the stream of a single-process run, a rank only touches the groupings it
                # owns); a single process draws in _filter_one_group, in the background of the upload.
                possible_probes = as_lists[group_i] if prefetch is not None else input[group_i]
                if failure is not None:
                    if prefetch is not None:
                        prefetch.get(group_i)
                        prefetch.release(k)
                    continue                     # keep taking part in the collectives below, then raise
                try:
                    pre = None
                    if prefetch is not None:
                        pre = prefetch.get(group_i) + (lambda k=k: prefetch.release(k),)
                    local[group_i] = self._filter_one_group(group_i, len(input), possible_probes,
                                                            target_genomes_grouped[group_i], target_genomes_grouped,
                                                            chain, pre)
                except Exception as e:
                    if not sharded:
                        raise
                    failure = e
        finally:
            if prefetch is not None:
                prefetch.close()
        if sharded:
            chain.finish()
        chosen_per_group = parallel.exchange_group_results(local, owner, rank, failure,
                                                           token=cov.fingerprint_lists(input)) if sharded else \
            [local[i] for i in range(len(input))]
        selected = []
        for possible_probes, chosen in zip(input, chosen_per_group):
            if not isinstance(possible_probes, (list, tuple, ProbeBatch)):
                possible_probes = list(possible_probes)
            selected.append(possible_probes.probes(list(chosen)) if isinstance(possible_probes, ProbeBatch) else
                            [possible_probes[i] for i in chosen])
        return selected

    def _run_pipelined(self, input, target_genomes_grouped, mine, chain, n_workers):
        """The groupings `mine` on n_workers threads, each driving its own context (stream, staging buffers) on the
        device: while one grouping is in a library call (the GIL is released), the host work of the next -- gather,
        seed plan, result lists -- runs on the other thread, and its uploads and kernels queue up behind on their own
        stream.  The seed draws come from ONE helper thread in grouping order (a single process replays the stream
        exactly like a rank of a group-sharded run does, _DrawReplay), so the workers never touch numpy's RNG.
        Returns ({grouping: selected indices}, first exception or None)."""
        import queue
        own_chain = chain is None
        if own_chain:
            chain = _DrawReplay(self, input, [0] * len(input), 0)
        main = self._context()
        contexts = [main] + [_lib.extra_context(main.device_id, w) for w in range(1, n_workers)]
        todo = queue.SimpleQueue()
        for g in mine:
            todo.put(g)
        local, errors = {}, []

        def work(ctx):
            self._tls.ctx = ctx
            try:
                while not errors:
                    try:
                        g = todo.get_nowait()
                    except queue.Empty:
                        return
                    probes = input[g] if isinstance(input[g], (list, tuple, ProbeBatch)) else list(input[g])
                    local[g] = self._filter_one_group(g, len(input), probes, target_genomes_grouped[g],
                                                      target_genomes_grouped, chain, None)
            except BaseException as e:               # noqa: BLE001 -- handed to the caller
                errors.append(e)
            finally:
                self._tls.ctx = None

        threads = [threading.Thread(target=work, args=(c,), name='cb-group-worker') for c in contexts[1:]]
        for t in threads:
            t.start()
        work(main)
        for t in threads:
            t.join()
        if own_chain:
            chain.finish()
        for e in errors:
            if isinstance(e, _LengthsNotUniform):
                raise e
        return local, (errors[0] if errors else None)

    def _filter_one_group(self, group_i, n_groups, possible_probes, target_genomes, target_genomes_grouped, chain,
                          pre=None):
        """One grouping this process owns: gather, upload, seed plan, both stages.  `chain` is the draw
        chain of a group-sharded run (None in a single process, which draws here in the background of
        the upload).  `pre`: (gathered probes, staged targets, release callback) when the prefetch thread
        already gathered this grouping.  Returns the indices of the selected probes in the reference's
        output order."""
        sharded = chain is not None
        if not isinstance(possible_probes, (list, tuple, ProbeBatch)):
            possible_probes = list(possible_probes)
        n_probes = len(possible_probes)
        probe_strs = _LazyStrs(possible_probes)
        group = None
        lengths, dups = None, True
        host_ms = {}
        t_mark = time.perf_counter()

        def mark(name):
            nonlocal t_mark
            now = time.perf_counter()
            host_ms[name] = host_ms.get(name, 0.0) + (now - t_mark) * 1e3
            t_mark = now
        drawn = drawn_tol = None
        guess = None
        gathered = staged = release = None
        if pre is not None:
            gathered, staged, release = pre
        try:
            if sharded:
                pass                             # the draws come from the chain / replay thread, fetched after the upload
            elif n_probes and gathered is not None:
                # prefetched: the lengths are known, the draw runs on the library's worker thread during the upload
                drawn = cov.draw_seeds(gathered[1], self.mismatches, self.lcf_thres, self.kmer_probe_map_k,
                                       background=True)
            elif n_probes:
                # The seed draw needs only the probe lengths and runs on the library's worker thread
                # while the sequences are gathered, copied to the device and packed.  It is started
                # on the guess that all probes are as long as the first (candidate probes are);
                # if the gathered lengths say otherwise the guess is dropped -- a background draw
                # only touches numpy's RNG state when it is accepted -- and the draw is redone.
                guess = possible_probes.probe_length if isinstance(possible_probes, ProbeBatch) else \
                    len(possible_probes[0].seq_str)
                try:
                    drawn = cov.draw_seeds(np.full(n_probes, guess, dtype=np.int32), self.mismatches,
                                           self.lcf_thres, self.kmer_probe_map_k, background=True)
                except ValueError:          # e.g. k longer than the first probe: decided on the real lengths
                    drawn = None
            if n_probes:
                if gathered is None:
                    # sequences of the whole list in one buffer (straight into page-locked staging memory
                    # when this rank is going to upload them)
                    try:
                        gathered = cov.gather_staged(self._context(), 0, possible_probes)
                        if gathered is None:
                            gathered = cov.gather_probes(possible_probes)
                    except BaseException:
                        if drawn is not None:
                            cov.cancel_draw(drawn)
                        raise
                    lengths = gathered[1]
                    if not sharded and (drawn is None or not bool(np.all(lengths == guess))):
                        if drawn is not None:
                            cov.cancel_draw(drawn)
                        drawn = cov.draw_seeds(lengths, self.mismatches, self.lcf_thres, self.kmer_probe_map_k,
                                               background=True)
                lengths = gathered[1]
                mark('gather')
            if n_probes:
                # The tolerant draw (ranks) continues the same stream, so it follows once the first
                # has finished.
                try:
                    group = cov.PackedGroup(self._context(), possible_probes, target_genomes, gathered=gathered,
                                            targets_staged=staged)
                    mark('pack_and_upload')
                    dups = self._context().probes_have_duplicates(group.probes)
                    mark('duplicate_check')
                finally:
                    if not sharded:
                        drawn = cov.finish_draw(drawn)
                if sharded:
                    # only now: gather, upload and packing do not need the draws, and the wait for this grouping's
                    # turn in the stream overlaps them
                    drawn, drawn_tol = chain.get(group_i, lengths)
                if self._needs_ranks() and not sharded:
                    drawn_tol = cov.draw_seeds(lengths, self.mismatches_tolerant, self.lcf_thres_tolerant,
                                               self.kmer_probe_map_k)
                mark('seed_draw_wait')
            elif sharded:
                chain.get(group_i)
        finally:
            if release is not None:
                release()                        # the staging pair may take the next grouping
        plan = plan_tol = None
        if n_probes:
            plan = cov.SeedPlan(probe_strs, self.mismatches, self.lcf_thres, self.kmer_probe_map_k,
                                lengths=lengths, may_have_dups=dups, drawn=drawn)
            if self._needs_ranks():
                plan_tol = cov.SeedPlan(probe_strs, self.mismatches_tolerant, self.lcf_thres_tolerant,
                                        self.kmer_probe_map_k, lengths=lengths, may_have_dups=dups,
                                        drawn=drawn_tol)
            mark('seed_plan')
        self._tls.host_ms = host_ms
        return self._select_for_group(group_i, n_groups, probe_strs, group, plan, plan_tol,
                                      target_genomes, target_genomes_grouped)

    def _shard_probes(self, n_groups, world_size):
        mode = os.environ.get('CB_SHARD', 'auto')
        if mode == 'groups' or self._needs_ranks():
            return False
        return mode == 'probes' or n_groups < world_size

    def _filter_probe_sharded(self, input, target_genomes_grouped, rank, world_size):
        """Every rank works on every grouping: stage A on its block of the probes, then the
        greedy loop with gains and interval index sharded and one exchange per round between the GPUs
        (cb_setcover_sharded); every rank ends up with the same picks.  Groupings that need p_u < 1 take
        the all-gather of the coverage and a replicated greedy instead."""
        ctx = self._context()
        selected = []
        for group_i, (possible_probes, target_genomes) in enumerate(zip(input, target_genomes_grouped)):
            if not isinstance(possible_probes, (list, tuple, ProbeBatch)):
                possible_probes = list(possible_probes)
            n_probes = len(possible_probes)
            t0 = time.perf_counter()
            stats = {'group': group_i, 'n_probes': n_probes, 'shard': 'probes',
                     'target_bp': sum(g.size() for g in target_genomes)}
            self.last_stats[group_i] = stats
            if n_probes == 0:
                selected.append([])
                continue
            lo, hi = parallel.shard_bounds(n_probes, world_size, rank)
            token = parallel.rng_state_token()
            local = group = None
            failure = None
            universe_p = self._make_universe_p(target_genomes)
            sharded_b = bool(np.all(universe_p == 1.0))
            try:
                # the draw consumes the RNG for ALL probes on every rank (same stream everywhere, and the
                # stream ends where a single process would leave it); each rank uses its own rows
                guess = possible_probes.probe_length if isinstance(possible_probes, ProbeBatch) else \
                    len(possible_probes[0].seq_str)
                try:
                    drawn = cov.draw_seeds(np.full(n_probes, guess, dtype=np.int32), self.mismatches,
                                           self.lcf_thres, self.kmer_probe_map_k, background=True)
                except ValueError:
                    drawn = None
                try:
                    gathered = cov.gather_staged(ctx, 0, possible_probes)
                    if gathered is None:
                        gathered = cov.gather_probes(possible_probes)
                except BaseException:
                    if drawn is not None:
                        cov.cancel_draw(drawn)
                    raise
                lengths = gathered[1]
                # every rank must hold the SAME list in the same order (it shards it by position)
                token = (token * 1000003) ^ cov.fingerprint(gathered)
                if drawn is None or not bool(np.all(lengths == guess)):
                    if drawn is not None:
                        cov.cancel_draw(drawn)
                    drawn = cov.draw_seeds(lengths, self.mismatches, self.lcf_thres, self.kmer_probe_map_k,
                                           background=True)
                try:
                    group = cov.PackedGroup(ctx, possible_probes, target_genomes, gathered=gathered)
                    dups = ctx.probes_have_duplicates(group.probes)
                finally:
                    drawn = cov.finish_draw(drawn)
                plan = cov.SeedPlan(_LazyStrs(possible_probes), self.mismatches, self.lcf_thres,
                                    self.kmer_probe_map_k, lengths=lengths, may_have_dups=dups, drawn=drawn)
                local, st_a = cov.compute_cover_range(ctx, group, plan, self.mismatches, self.lcf_thres,
                                                      self.island_of_exact_match, self.cover_extension, lo, hi)
            except Exception as e:
                failure = e
            finally:
                if group is not None:
                    group.free()
            try:
                # collective from here on; a rank that failed above says so and everybody raises
                need = ctx.exchange_required(local) if (failure is None and sharded_b) else 0
                try:
                    parallel.ensure_exchange(ctx, need, token, failed=failure is not None)
                except RuntimeError:
                    if failure is not None:
                        raise failure
                    raise
                if sharded_b:
                    picks, st_b = ctx.setcover_sharded(local, n_probes, lo, hi)
                else:
                    parallel.ensure_comm(ctx)
                    cover = ctx.cover_allgather(local, lo, n_probes)
                    try:
                        picks, st_b = ctx.setcover(cover, n_probes, None, universe_p)
                    finally:
                        cover.free()
            finally:
                if local is not None:
                    local.free()
            chosen = set()
            for i in picks.tolist():
                chosen.add(i)
            chosen = pickle.loads(pickle.dumps(chosen))
            selected.append(possible_probes.probes(list(chosen)) if isinstance(possible_probes, ProbeBatch) else
                            [possible_probes[i] for i in chosen])
            seed_bytes = int(plan.uniform[lo:hi].nbytes) if plan.uniform is not None else \
                int(plan.seed_pos.nbytes + plan.seed_off.nbytes)
            stats.update(seed_mode=plan.mode, k=plan.k, bits=group.bits, h2d_bytes=group.h2d_bytes + seed_bytes,
                         d2h_bytes=int(picks.nbytes), picks=picks, upload_targets=group.st_targets.as_dict(),
                         upload_probes=group.st_probes.as_dict(), coverage=st_a.as_dict(),
                         setcover=st_b.as_dict(), wall_s=time.perf_counter() - t0)
        return selected

    def _select_for_group(self, group_i, n_groups, probe_strs, group, plan, plan_tol, target_genomes,
                          target_genomes_grouped):
        """Indices (into the grouping's probe list) of the selected probes, in the reference's
        output order."""
        ctx = self._context()
        t0 = time.perf_counter()
        stats = {'group': group_i, 'n_probes': len(probe_strs),
                 'target_bp': group.target_bases if group is not None else sum(g.size() for g in target_genomes)}
        self.last_stats[group_i] = stats
        if len(probe_strs) == 0:                            # set_cover_filter.py:393-394
            return []
        logger.info("Computing coverage of %d probes in %d genomes (group %d of %d)",
                    len(probe_strs), len(target_genomes), group_i + 1, n_groups)
        try:
            cover, st_a = cov.compute_cover(ctx, group, plan, self.mismatches, self.lcf_thres,
                                            self.island_of_exact_match, self.cover_extension)
        finally:
            group.free()
        try:
            ranks = self._make_ranks(ctx, probe_strs, plan_tol, target_genomes_grouped) \
                if plan_tol is not None else None
            universe_p = self._make_universe_p(target_genomes)
            picks, st_b = ctx.setcover(cover, len(probe_strs), ranks, universe_p)
        finally:
            cover.free()
        if ranks is not None:
            n_bad = int(np.sum(ranks[picks] > 0))
            if n_bad > 0:
                logger.warning("Group %d: forced to choose %d less-than-ideal probe%s",
                               group_i + 1, n_bad, '' if n_bad == 1 else 's')
        # The reference returns a Python set of ids from a Pool worker and iterates it
        # (set_cover_filter.py:893-900, :926): same elements, CPython set order after a
        # pickle round trip.  Reproduce it with the real thing.
        chosen = set()
        for i in picks.tolist():
            chosen.add(i)
        chosen = pickle.loads(pickle.dumps(chosen))
        seed_bytes = int(plan.uniform.nbytes) if plan.uniform is not None else \
            int(plan.seed_pos.nbytes + plan.seed_off.nbytes)
        stats.update(seed_mode=plan.mode, k=plan.k, bits=group.bits, h2d_bytes=group.h2d_bytes + seed_bytes,
                     d2h_bytes=int(picks.nbytes), picks=picks,
                     upload_targets=group.st_targets.as_dict(),
                     upload_probes=group.st_probes.as_dict(),
                     coverage=st_a.as_dict(), setcover=st_b.as_dict(),
                     wall_s=time.perf_counter() - t0, host_ms=dict(getattr(self._tls, 'host_ms', {})))
        return list(chosen)
