"""ctypes binding of libcatchb200.so (include/catch_b200.h).  Fails loudly when the library is
missing or when a call returns an error -- there is no CPU fallback on the product path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcatchb200.so')


class CatchB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcatchb200 error %d: %s" % (code, msg))
        self.code = code


class HybParams(C.Structure):
    _fields_ = [('mismatches', C.c_int32), ('lcf_thres', C.c_int32),
                ('island_of_exact_match', C.c_int32), ('cover_extension', C.c_int32),
                ('k', C.c_int32)]


class Stats(C.Structure):
    _fields_ = [('ms_h2d', C.c_double), ('ms_pack', C.c_double), ('ms_seed_index', C.c_double),
                ('ms_scan_count', C.c_double), ('ms_scan_emit', C.c_double), ('ms_merge', C.c_double),
                ('ms_universe', C.c_double), ('ms_greedy', C.c_double), ('ms_d2h', C.c_double),
                ('ms_total', C.c_double),
                ('n_seed_entries', C.c_int64), ('n_seed_lookups', C.c_int64),
                ('n_candidate_hits', C.c_int64), ('n_raw_ranges', C.c_int64),
                ('n_intervals', C.c_int64), ('n_picks', C.c_int64), ('n_kernel_launches', C.c_int64),
                ('bytes_algorithmic', C.c_int64), ('reserved', C.c_int64 * 8)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != 'reserved'}
        d['reserved'] = list(self.reserved)
        return d


# every symbol include/catch_b200.h declares
EXPORTED_SYMBOLS = [
    'cb_init', 'cb_destroy', 'cb_last_error', 'cb_version', 'cb_flush_l2', 'cb_intop_rate', 'cb_pool_reserve', 'cb_host_buffer',
    'cb_upload_targets', 'cb_targets_free', 'cb_upload_probes', 'cb_probes_free', 'cb_upload_group',
    'cb_probes_have_duplicates', 'cb_mt19937_randint', 'cb_mt19937_randint_scalar', 'cb_mt19937_randint_u8',
    'cb_mt19937_randint_begin',
    'cb_mt19937_randint_end',
    'cb_split_lengths',
    'cb_coverage', 'cb_coverage_uniform', 'cb_cover_free', 'cb_cover_num_intervals', 'cb_cover_export', 'cb_cover_import',
    'cb_setcover', 'cb_setcover_costs', 'cb_minhash_neardup', 'cb_hamming_neardup', 'cb_neardup_filter', 'cb_group_duplicates',
    'cb_comm_unique_id', 'cb_comm_init', 'cb_comm_destroy', 'cb_cover_allgather',
    'cb_coverage_range', 'cb_coverage_records', 'cb_free_host', 'cb_exchange_alloc', 'cb_exchange_bytes', 'cb_exchange_handle', 'cb_exchange_attach',
    'cb_exchange_required', 'cb_setcover_sharded', 'cb_setcover_sharded_begin', 'cb_setcover_sharded_end',
    'cb_sketch_sequences', 'cb_sketches_import', 'cb_sketches_export', 'cb_sketches_free', 'cb_sketch_dist_rows',
    'cb_sketch_dist_condensed', 'cb_sketch_near_rows',
]

_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "catch_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C catch_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.cb_version.restype = C.c_char_p
    L.cb_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.cb_destroy.argtypes = [vp]
    L.cb_destroy.restype = None
    L.cb_last_error.argtypes = [vp]
    L.cb_last_error.restype = C.c_char_p
    L.cb_flush_l2.argtypes = [vp]
    L.cb_intop_rate.argtypes = [vp, C.POINTER(C.c_double)]
    L.cb_pool_reserve.argtypes = [vp, i64]
    L.cb_host_buffer.argtypes = [vp, i32, i64, C.POINTER(vp)]
    L.cb_upload_targets.argtypes = [vp, vp, vp, i64, vp, i32, vp, i32, C.POINTER(vp), C.POINTER(Stats)]
    L.cb_targets_free.argtypes = [vp]
    L.cb_targets_free.restype = None
    L.cb_upload_probes.argtypes = [vp, vp, vp, i64, vp, i32, C.POINTER(vp), C.POINTER(Stats)]
    L.cb_probes_free.argtypes = [vp]
    L.cb_upload_group.argtypes = [vp, vp, i64, vp, i64, i32, vp, vp, i64, vp, i32, vp, C.POINTER(i32),
                                  C.POINTER(vp), C.POINTER(vp), C.POINTER(Stats)]
    L.cb_probes_free.restype = None
    L.cb_probes_have_duplicates.argtypes = [vp, vp, C.POINTER(i32)]
    L.cb_mt19937_randint_scalar.argtypes = [vp, C.POINTER(i32), C.c_uint32, i64, vp]
    L.cb_mt19937_randint_u8.argtypes = [vp, C.POINTER(i32), C.c_uint32, i64, vp]
    L.cb_mt19937_randint_begin.argtypes = [vp, C.POINTER(i32), C.c_uint32, i64, vp, i32]
    L.cb_mt19937_randint_begin.restype = vp
    L.cb_mt19937_randint_end.argtypes = [vp]
    L.cb_split_lengths.argtypes = [vp, i64, i64, i32, vp]
    L.cb_mt19937_randint.argtypes = [vp, C.POINTER(i32), C.c_uint32, i64, vp]
    L.cb_coverage.argtypes = [vp, vp, vp, C.POINTER(HybParams), vp, vp, C.POINTER(vp), C.POINTER(Stats)]
    L.cb_coverage_uniform.argtypes = [vp, vp, vp, C.POINTER(HybParams), vp, i32, C.POINTER(vp), C.POINTER(Stats)]
    L.cb_coverage_range.argtypes = [vp, vp, vp, C.POINTER(HybParams), vp, vp, vp, i32, i64, i64, C.POINTER(vp),
                                    C.POINTER(Stats)]
    L.cb_coverage_records.argtypes = [vp, vp, vp, C.POINTER(HybParams), vp, vp, C.POINTER(i64), C.POINTER(vp),
                                      C.POINTER(Stats)]
    L.cb_free_host.argtypes = [vp]
    L.cb_free_host.restype = None
    L.cb_exchange_alloc.argtypes = [vp, i64]
    L.cb_exchange_bytes.argtypes = [vp]
    L.cb_exchange_bytes.restype = i64
    L.cb_exchange_handle.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.cb_exchange_attach.argtypes = [vp, i32, i32, vp, vp, i32]
    L.cb_exchange_required.argtypes = [vp, vp, C.POINTER(i64)]
    L.cb_setcover_sharded.argtypes = [vp, vp, i64, i64, vp, vp, C.POINTER(i64), C.POINTER(Stats)]
    L.cb_setcover_sharded_begin.argtypes = [vp, vp, i64, i64, vp, C.POINTER(vp)]
    L.cb_setcover_sharded_end.argtypes = [vp, vp, vp, C.POINTER(i64), C.POINTER(Stats)]
    L.cb_cover_free.argtypes = [vp]
    L.cb_cover_free.restype = None
    L.cb_cover_num_intervals.argtypes = [vp]
    L.cb_cover_num_intervals.restype = i64
    L.cb_cover_export.argtypes = [vp, vp, vp, vp, vp, vp]
    L.cb_cover_import.argtypes = [vp, i64, i32, vp, i64, vp, vp, vp, vp, C.POINTER(vp)]
    L.cb_comm_unique_id.argtypes = [vp, vp]
    L.cb_comm_init.argtypes = [vp, vp, i32, i32]
    L.cb_comm_destroy.argtypes = [vp]
    L.cb_cover_allgather.argtypes = [vp, vp, i64, i64, C.POINTER(vp)]
    L.cb_setcover.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(Stats)]
    L.cb_setcover_costs.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(Stats)]
    L.cb_minhash_neardup.argtypes = [vp, vp, vp, i64, vp, vp, i32, i32, i32, C.c_double, vp, C.POINTER(Stats)]
    L.cb_neardup_filter.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, i32, i32, i32, C.c_double, vp, C.POINTER(i64),
                                    C.POINTER(i64), C.POINTER(Stats)]
    L.cb_group_duplicates.argtypes = [vp, vp, vp, i64, vp, vp, C.POINTER(i64), C.POINTER(Stats)]
    L.cb_hamming_neardup.argtypes = [vp, vp, vp, i64, vp, i32, i32, i32, vp, C.POINTER(Stats)]
    L.cb_sketch_sequences.argtypes = [vp, vp, vp, i64, i32, i32, C.c_uint64, C.c_uint64, C.POINTER(vp), C.POINTER(Stats)]
    L.cb_sketches_import.argtypes = [vp, vp, i64, i32, C.POINTER(vp)]
    L.cb_sketches_export.argtypes = [vp, vp, vp]
    L.cb_sketches_free.argtypes = [vp]
    L.cb_sketches_free.restype = None
    L.cb_sketch_dist_rows.argtypes = [vp, vp, vp, i64, vp]
    L.cb_sketch_dist_condensed.argtypes = [vp, vp, vp]
    L.cb_sketch_near_rows.argtypes = [vp, vp, vp, i64, C.c_double, vp, C.POINTER(vp), C.POINTER(vp)]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data if a is not None else None


_ONE_BYTE = np.zeros(1, dtype=np.uint8)


class Context:
    """One cb_ctx (one CUDA device, one stream).  Not thread-safe, as in the reference where a
    filter instance is not re-entrant (probe.py:820-832)."""

    def __init__(self, device_id=None):
        self.L = load()
        if device_id is None:
            device_id = int(os.environ.get('LOCAL_RANK', os.environ.get('CB_DEVICE', '0')))
        self.device_id = device_id
        self._pinned_addr, self._pinned_cap = {}, {}
        h = C.c_void_p()
        rc = self.L.cb_init(device_id, C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.L.cb_last_error(h).decode() if h else 'cb_init failed'
            raise CatchB200Error(rc, msg)

    def close(self):
        if getattr(self, 'h', None):
            self.L.cb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CatchB200Error(rc, self.L.cb_last_error(self.h).decode())

    def host_buffer(self, slot, nbytes):
        """(address, capacity) of the context's page-locked staging buffer `slot`, grown to hold
        at least nbytes."""
        cap = self._pinned_cap.get(slot, 0)
        if nbytes > cap or slot not in self._pinned_addr:
            out = C.c_void_p()
            self._check(self.L.cb_host_buffer(self.h, slot, max(int(nbytes), 1), C.byref(out)))
            self._pinned_addr[slot] = out.value
            self._pinned_cap[slot] = max(int(nbytes), 1)
        return self._pinned_addr[slot], self._pinned_cap[slot]

    def flush_l2(self):
        self._check(self.L.cb_flush_l2(self.h))

    def pool_reserve(self, nbytes):
        self._check(self.L.cb_pool_reserve(self.h, int(nbytes)))

    def intop_rate(self):
        """Measured integer-ALU rate of the device (ops/s), see cb_intop_rate."""
        v = C.c_double()
        self._check(self.L.cb_intop_rate(self.h, C.byref(v)))
        return float(v.value)

    # ---- packing
    def upload_targets(self, ascii_u8, seq_off, seq_genome, n_genomes, lut, bits):
        out, st = C.c_void_p(), Stats()
        self._check(self.L.cb_upload_targets(self.h, _ptr(ascii_u8), _ptr(seq_off), len(seq_off) - 1,
                                             _ptr(seq_genome), n_genomes, _ptr(lut), bits,
                                             C.byref(out), C.byref(st)))
        return Handle(self.L.cb_targets_free, out), st

    def upload_probes(self, ascii_u8, probe_off, lut, bits):
        out, st = C.c_void_p(), Stats()
        self._check(self.L.cb_upload_probes(self.h, _ptr(ascii_u8), _ptr(probe_off), len(probe_off) - 1,
                                            _ptr(lut), bits, C.byref(out), C.byref(st)))
        return Handle(self.L.cb_probes_free, out), st

    def upload_group(self, probes_raw, n_probes, targets_raw, seq_off, seq_genome, n_genomes, probe_off=None,
                     sep=10):
        """cb_upload_group: `probes_raw` is a bytes object holding the probes separated by `sep`
        (or back to back with explicit `probe_off`).  Returns (probes, targets, lengths, bits, stats)."""
        p_out, t_out, st, bits = C.c_void_p(), C.c_void_p(), Stats(), C.c_int32()
        lens = np.zeros(max(n_probes, 1), dtype=np.int32)
        # probes_raw / targets_raw: bytes objects, or integer addresses of staged host memory
        # (then probe_off gives the size)
        p_bytes = len(probes_raw) if isinstance(probes_raw, (bytes, bytearray)) else int(probe_off[-1])
        self._check(self.L.cb_upload_group(self.h, probes_raw, p_bytes, _ptr(probe_off), n_probes, sep,
                                           targets_raw, _ptr(seq_off), len(seq_off) - 1, _ptr(seq_genome), n_genomes,
                                           _ptr(lens), C.byref(bits), C.byref(p_out), C.byref(t_out), C.byref(st)))
        return (Handle(self.L.cb_probes_free, p_out), Handle(self.L.cb_targets_free, t_out), lens[:n_probes],
                bits.value, st)

    def probes_have_duplicates(self, probes):
        flag = C.c_int32()
        self._check(self.L.cb_probes_have_duplicates(self.h, probes.h, C.byref(flag)))
        return bool(flag.value)

    # ---- stage A
    def coverage(self, probes, targets, mismatches, lcf_thres, island, cover_extension, k,
                 seed_off, seed_pos):
        hp = HybParams(mismatches, lcf_thres, island, cover_extension, k)
        out, st = C.c_void_p(), Stats()
        self._check(self.L.cb_coverage(self.h, probes.h, targets.h, C.byref(hp), _ptr(seed_off),
                                       _ptr(seed_pos), C.byref(out), C.byref(st)))
        return Handle(self.L.cb_cover_free, out), st

    def coverage_uniform(self, probes, targets, mismatches, lcf_thres, island, cover_extension, k, seeds_u8):
        """cb_coverage_uniform: `seeds_u8` is a C-contiguous uint8 [n_probes, s] matrix."""
        hp = HybParams(mismatches, lcf_thres, island, cover_extension, k)
        out, st = C.c_void_p(), Stats()
        s = seeds_u8.shape[1] if seeds_u8.ndim == 2 else 0
        self._check(self.L.cb_coverage_uniform(self.h, probes.h, targets.h, C.byref(hp),
                                               seeds_u8.ctypes.data if seeds_u8.size else _ONE_BYTE.ctypes.data, s,
                                               C.byref(out), C.byref(st)))
        return Handle(self.L.cb_cover_free, out), st

    def coverage_range(self, probes, targets, mismatches, lcf_thres, island, cover_extension, k, lo, hi,
                       seeds_u8=None, seed_off=None, seed_pos=None):
        """cb_coverage_range: stage A for the probes [lo, hi) only (a rank's shard); the cover keeps global
        probe ids.  `seeds_u8`: uint8 [hi - lo, s] matrix, first row for probe lo; or the CSR pair."""
        hp = HybParams(mismatches, lcf_thres, island, cover_extension, k)
        out, st = C.c_void_p(), Stats()
        if seeds_u8 is not None:
            seeds_u8 = np.ascontiguousarray(seeds_u8, dtype=np.uint8)
            s = seeds_u8.shape[1] if seeds_u8.ndim == 2 else 0
            self._check(self.L.cb_coverage_range(self.h, probes.h, targets.h, C.byref(hp), None, None,
                                                 seeds_u8.ctypes.data if seeds_u8.size else _ONE_BYTE.ctypes.data, s,
                                                 lo, hi, C.byref(out), C.byref(st)))
        else:
            self._check(self.L.cb_coverage_range(self.h, probes.h, targets.h, C.byref(hp), _ptr(seed_off),
                                                 _ptr(seed_pos), None, 0, lo, hi, C.byref(out), C.byref(st)))
        return Handle(self.L.cb_cover_free, out), st

    def coverage_records(self, probes, targets, mismatches, lcf_thres, island, k, seed_off, seed_pos):
        """cb_coverage_records: the unmerged ranges of the scan as an int64 [n, 5] array
        (probe, sequence, start, end, position of the seed hit), in no particular order."""
        hp = HybParams(mismatches, lcf_thres, island, 0, k)
        n, buf, st = C.c_int64(), C.c_void_p(), Stats()
        self._check(self.L.cb_coverage_records(self.h, probes.h, targets.h, C.byref(hp), _ptr(seed_off), _ptr(seed_pos),
                                               C.byref(n), C.byref(buf), C.byref(st)))
        try:
            if n.value == 0:
                return np.zeros((0, 5), dtype=np.int64), st
            raw = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint32)), shape=(n.value, 5))
            return raw.astype(np.int64), st
        finally:
            if buf.value:
                self.L.cb_free_host(buf)

    def cover_export(self, cover):
        n = self.L.cb_cover_num_intervals(cover.h)
        pid = np.zeros(n, dtype=np.int64)
        gen = np.zeros(n, dtype=np.int32)
        start = np.zeros(n, dtype=np.int64)
        end = np.zeros(n, dtype=np.int64)
        self._check(self.L.cb_cover_export(self.h, cover.h, _ptr(pid), _ptr(gen), _ptr(start), _ptr(end)))
        return pid, gen, start, end

    def cover_import(self, n_probes, genome_len, probe_id, genome, start, end):
        """Cover from host intervals (flat form of approx_multiuniverse's `sets`)."""
        gl = np.ascontiguousarray(genome_len, dtype=np.int64)
        pid = np.ascontiguousarray(probe_id, dtype=np.int64)
        gen = np.ascontiguousarray(genome, dtype=np.int32)
        s = np.ascontiguousarray(start, dtype=np.int64)
        e = np.ascontiguousarray(end, dtype=np.int64)
        out = C.c_void_p()
        self._check(self.L.cb_cover_import(self.h, n_probes, len(gl), _ptr(gl) if len(gl) else None, len(pid),
                                           _ptr(pid) if len(pid) else None, _ptr(gen) if len(pid) else None,
                                           _ptr(s) if len(pid) else None, _ptr(e) if len(pid) else None,
                                           C.byref(out)))
        return Handle(self.L.cb_cover_free, out)

    # ---- multi-GPU (probe sharding inside a grouping)
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._check(self.L.cb_comm_unique_id(self.h, buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.L.cb_comm_init(self.h, buf, rank, n_ranks))
        self.comm_ready = True

    def cover_allgather(self, local_cover, probe_lo, n_probes_total):
        out = C.c_void_p()
        self._check(self.L.cb_cover_allgather(self.h, local_cover.h, probe_lo, n_probes_total, C.byref(out)))
        return Handle(self.L.cb_cover_free, out)

    # ---- multi-GPU: sharded set cover through peer-mapped exchange areas
    def exchange_alloc(self, nbytes):
        self._check(self.L.cb_exchange_alloc(self.h, int(nbytes)))
        self.exchange_ready = False

    def exchange_bytes(self):
        return int(self.L.cb_exchange_bytes(self.h))

    def exchange_handle(self):
        """(64-byte CUDA IPC handle, device address) of this context's exchange area."""
        buf = (C.c_uint8 * 64)()
        addr = C.c_uint64()
        self._check(self.L.cb_exchange_handle(self.h, buf, C.byref(addr)))
        return bytes(buf), int(addr.value)

    def exchange_attach(self, rank, n_ranks, handles=None, addresses=None, grid_limit=0):
        hb = ab = None
        if handles is not None:
            hb = (C.c_uint8 * (64 * n_ranks)).from_buffer_copy(b''.join(handles))
        if addresses is not None:
            ab = (C.c_uint64 * n_ranks)(*addresses)
        self._check(self.L.cb_exchange_attach(self.h, rank, n_ranks, hb, ab, grid_limit))
        self.exchange_ready = True
        self.exchange_rank, self.exchange_n_ranks = rank, n_ranks

    def exchange_required(self, cover):
        n = C.c_int64()
        self._check(self.L.cb_exchange_required(self.h, cover.h, C.byref(n)))
        return int(n.value)

    def setcover_sharded(self, cover, n_probes, lo, hi, ranks=None):
        """cb_setcover_sharded (collective over the attached ranks): picks in pick order, the same on
        every rank."""
        sel = np.zeros(max(n_probes, 1), dtype=np.int64)
        n, st = C.c_int64(), Stats()
        self._check(self.L.cb_setcover_sharded(self.h, cover.h, lo, hi, _ptr(ranks), _ptr(sel), C.byref(n), C.byref(st)))
        return sel[:n.value].copy(), st

    def setcover_sharded_begin(self, cover, lo, hi, ranks=None):
        """cb_setcover_sharded_begin: the rank-local set-up; returns a job for setcover_sharded_end."""
        job = C.c_void_p()
        self._check(self.L.cb_setcover_sharded_begin(self.h, cover.h, lo, hi, _ptr(ranks), C.byref(job)))
        return job

    def setcover_sharded_end(self, job, n_probes):
        sel = np.zeros(max(n_probes, 1), dtype=np.int64)
        n, st = C.c_int64(), Stats()
        self._check(self.L.cb_setcover_sharded_end(self.h, job, _ptr(sel), C.byref(n), C.byref(st)))
        return sel[:n.value].copy(), st

    # ---- stage B
    def setcover(self, cover, n_probes, ranks=None, universe_p=None, costs=None):
        sel = np.zeros(max(n_probes, 1), dtype=np.int64)
        n, st = C.c_int64(), Stats()
        if costs is not None:
            costs = np.ascontiguousarray(costs, dtype=np.float64)
            self._check(self.L.cb_setcover_costs(self.h, cover.h, _ptr(costs), _ptr(ranks), _ptr(universe_p),
                                                 _ptr(sel), C.byref(n), C.byref(st)))
        else:
            self._check(self.L.cb_setcover(self.h, cover.h, _ptr(ranks), _ptr(universe_p), _ptr(sel),
                                           C.byref(n), C.byref(st)))
        return sel[:n.value].copy(), st

    # ---- near-duplicate filter
    def minhash_neardup(self, ascii_u8, probe_off, a, b, n_tables, k_concat, kmer_size, dist_thres):
        n = len(probe_off) - 1
        keep = np.zeros(max(n, 1), dtype=np.uint8)
        st = Stats()
        self._check(self.L.cb_minhash_neardup(self.h, _ptr(ascii_u8), _ptr(probe_off), n, _ptr(a), _ptr(b),
                                              n_tables, k_concat, kmer_size, float(dist_thres), _ptr(keep),
                                              C.byref(st)))
        return keep[:n], st

    def neardup_filter(self, raw, probe_off, family, a, b, positions, n_tables, k_concat, kmer_size, dist_thres):
        """cb_neardup_filter on the whole probe list (duplicates included).  `raw`: bytes object or the
        address of staged host memory holding the sequences back to back.  Returns (list indices of
        the first occurrence of every kept sequence in priority order, number of distinct sequences,
        stats)."""
        n = len(probe_off) - 1
        kept = np.zeros(max(n, 1), dtype=np.int64)
        nk, nd, st = C.c_int64(), C.c_int64(), Stats()
        self._check(self.L.cb_neardup_filter(self.h, raw, _ptr(probe_off), n, family, _ptr(a), _ptr(b),
                                             _ptr(positions), n_tables, k_concat, kmer_size, float(dist_thres),
                                             _ptr(kept), C.byref(nk), C.byref(nd), C.byref(st)))
        return kept[:nk.value].copy(), nd.value, st

    def group_duplicates(self, raw, probe_off):
        """cb_group_duplicates: (first-occurrence index, multiplicity) of every distinct sequence, in
        order of first occurrence."""
        n = len(probe_off) - 1
        first = np.zeros(max(n, 1), dtype=np.int64)
        count = np.zeros(max(n, 1), dtype=np.int32)
        nd, st = C.c_int64(), Stats()
        self._check(self.L.cb_group_duplicates(self.h, raw, _ptr(probe_off), n, _ptr(first), _ptr(count),
                                               C.byref(nd), C.byref(st)))
        return first[:nd.value], count[:nd.value], st

    def hamming_neardup(self, ascii_u8, probe_off, positions, n_tables, k_concat, dist_thres):
        n = len(probe_off) - 1
        keep = np.zeros(max(n, 1), dtype=np.uint8)
        st = Stats()
        self._check(self.L.cb_hamming_neardup(self.h, _ptr(ascii_u8), _ptr(probe_off), n, _ptr(positions),
                                              n_tables, k_concat, int(dist_thres), _ptr(keep), C.byref(st)))
        return keep[:n], st


    # ---- genome clustering: MinHash sketches
    def sketch_sequences(self, raw, seq_off, kmer_size, N, a, b):
        """cb_sketch_sequences: `raw` holds the sequences back to back (bytes), seq_off their int64 offsets.
        Returns (sketches handle, stats)."""
        out, st = C.c_void_p(), Stats()
        self._check(self.L.cb_sketch_sequences(self.h, raw, _ptr(seq_off), len(seq_off) - 1, kmer_size, N,
                                               int(a), int(b), C.byref(out), C.byref(st)))
        return Handle(self.L.cb_sketches_free, out), st

    def sketches_import(self, sig):
        out = C.c_void_p()
        self._check(self.L.cb_sketches_import(self.h, _ptr(sig), sig.shape[0], sig.shape[1], C.byref(out)))
        return Handle(self.L.cb_sketches_free, out)

    def sketches_export(self, sketches, n, N):
        sig = np.zeros((n, N), dtype=np.uint32)
        if n:
            self._check(self.L.cb_sketches_export(self.h, sketches.h, _ptr(sig)))
        return sig

    def sketch_dist_rows(self, sketches, rows, n):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros((len(rows), n), dtype=np.float64)
        if out.size:
            self._check(self.L.cb_sketch_dist_rows(self.h, sketches.h, _ptr(rows), len(rows), _ptr(out)))
        return out

    def sketch_near_rows(self, sketches, rows, threshold):
        """cb_sketch_near_rows: (row_off int64 [len(rows) + 1], columns uint32, distances float64) of the sketches
        within `threshold` of each row, columns ascending."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        off = np.zeros(len(rows) + 1, dtype=np.int64)
        idx, dist = C.c_void_p(), C.c_void_p()
        self._check(self.L.cb_sketch_near_rows(self.h, sketches.h, _ptr(rows), len(rows), float(threshold), _ptr(off),
                                               C.byref(idx), C.byref(dist)))
        total = int(off[-1])
        try:
            if total == 0:
                return off, np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.float64)
            return (off, np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_uint32)), shape=(total,)).copy(),
                    np.ctypeslib.as_array(C.cast(dist, C.POINTER(C.c_double)), shape=(total,)).copy())
        finally:
            if idx.value:
                self.L.cb_free_host(idx)
            if dist.value:
                self.L.cb_free_host(dist)

    def sketch_dist_condensed(self, sketches, n):
        out = np.zeros(n * (n - 1) // 2, dtype=np.float32)
        if out.size:
            self._check(self.L.cb_sketch_dist_condensed(self.h, sketches.h, _ptr(out)))
        return out


class Handle:
    """Owns an opaque device object and frees it with the matching cb_*_free."""

    def __init__(self, free_fn, h):
        self._free = free_fn
        self.h = h

    def free(self):
        if self.h:
            self._free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_default_ctx = None


def default_context():
    """Process-wide context, created on first use (never at import time, so forking before the
    first filter call stays safe)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


_extra_ctx = {}


def extra_context(device_id, index):
    """Further contexts (own stream, own staging buffers) on a device, created on first use and kept: a filter that
    works on two groupings at a time gives each worker thread its own."""
    key = (int(device_id), int(index))
    if key not in _extra_ctx:
        _extra_ctx[key] = Context(device_id)
    return _extra_ctx[key]


def split_lengths(raw, n, sep=10):
    """Lengths of the n strings joined with `sep` into the bytes object `raw`; None when the
    separator also occurs inside a string."""
    lens = np.zeros(max(n, 1), dtype=np.int32)
    rc = load().cb_split_lengths(raw, len(raw), n, sep, lens.ctypes.data)
    return lens[:n] if rc == 0 else None


def _small(bound):
    return np.uint8 if bound <= 256 else np.int32


def legacy_randint(bound, shape):
    """np.random.randint(0, bound, size=shape) on numpy's legacy global stream, generated by the
    library's MT19937 replay (same values, same final state, a fraction of the time).  Values come
    back as uint8 when bound <= 256 (seed positions), else int32."""
    L = load()
    name, key, pos, has_gauss, cached = np.random.get_state()
    dt = _small(bound)
    if name != 'MT19937':
        return np.random.randint(0, bound, size=shape).astype(dt)
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    n = int(np.prod(shape))
    out = np.empty(n, dtype=dt)
    p = C.c_int32(int(pos))
    fn = L.cb_mt19937_randint_u8 if dt is np.uint8 else L.cb_mt19937_randint
    rc = fn(key.ctypes.data, C.byref(p), int(bound), n, out.ctypes.data)
    if rc != 0:
        raise CatchB200Error(rc, 'cb_mt19937_randint')
    np.random.set_state((name, key, p.value, has_gauss, cached))
    return out.reshape(shape)


class PendingRandint:
    """legacy_randint running on the library's worker thread: result() waits for it, advances
    numpy's global state and returns the draws.  Nothing else may use np.random in between."""

    def __init__(self, bound, shape):
        self.L = load()
        self.shape = shape
        st = np.random.get_state()
        self.sync = None
        dt = _small(bound)
        if st[0] != 'MT19937':
            self.sync = np.random.randint(0, bound, size=shape).astype(dt)
            return
        self.name, key, pos, self.has_gauss, self.cached = st
        self.key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        self.out = np.empty(int(np.prod(shape)), dtype=dt)
        self.pos = C.c_int32(int(pos))
        self.job = self.L.cb_mt19937_randint_begin(self.key.ctypes.data, C.byref(self.pos), int(bound),
                                                   self.out.size, self.out.ctypes.data, self.out.itemsize)

    def cancel(self):
        """Wait for the worker and drop its output; numpy's state stays as it was."""
        if self.sync is None and self.job is not None:
            self.L.cb_mt19937_randint_end(self.job)
            self.job = None

    def result(self):
        if self.sync is not None:
            return self.sync
        rc = self.L.cb_mt19937_randint_end(self.job)
        self.job = None
        if rc != 0:
            raise CatchB200Error(rc, 'cb_mt19937_randint')
        np.random.set_state((self.name, self.key, self.pos.value, self.has_gauss, self.cached))
        return self.out.reshape(self.shape)
