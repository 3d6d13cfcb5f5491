/* _fastpack: host-side glue (CPython C API) that gathers the sequences of a list of Probe
 * objects into one contiguous byte buffer plus a length array, in a single pass and without
 * creating intermediate Python objects.  It is the C counterpart of
 *     [p.seq_str for p in probes]; ''.join(...).encode('latin-1'); lengths
 * which dominates the host time of SetCoverFilter.filter() on ~10^5 probes.  No CUDA here; the
 * buffers go to libcatchb200.so (cb_upload_group) through ctypes. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <descrobject.h>
#include <structmember.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pass 2 for large inputs: the copies are plain memcpy from buffers we hold references to, so they
 * run on a few threads with the GIL released (68 MB of genome segments: ~15 ms -> ~3 ms) */
typedef struct {
    const void **src;
    const int32_t *len;
    char *dst;
    Py_ssize_t begin, end;
} copy_job;

static void *copy_worker(void *arg)
{
    copy_job *j = (copy_job *)arg;
    char *d = j->dst;
    for (Py_ssize_t i = j->begin; i < j->end; i++) {
        if (i + 8 < j->end) __builtin_prefetch(j->src[i + 8]);      /* the strings are scattered over the heap */
        memcpy(d, j->src[i], (size_t)j->len[i]);
        d += j->len[i];
    }
    return NULL;
}

/* copies of at least this many bytes are spread over PAR_COPY_THREADS threads (CB_GATHER_PAR_MB overrides).  Measured on
 * the B200 host (tools/gather_sweep.py, 100-byte strings): 7 MB 1.05 ms on one thread / 1.87 ms on six, 13 MB 2.2 / 2.8,
 * 26 MB 6.1 / 5.3, 52 MB 27.8 / 17.3, 133 MB 73 / 48: the threads pay from about 20 MB on. */
static size_t par_copy_min_bytes(void)
{
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("CB_GATHER_PAR_MB");
        long mb = e ? atol(e) : 24;
        v = (size_t)(mb > 0 ? mb : 24) << 20;
    }
    return v;
}
#define PAR_COPY_THREADS 6

/* gather(seq, attr) -> (bytes data, bytes lengths_int32)
 * gather_into(seq, attr, address, capacity) -> (int total_bytes, bytes lengths_int32)
 * Every item is either a str or an object whose attribute `attr` is a str.  Raises ValueError
 * when a string holds characters above U+00FF (not representable as one byte per base).
 * gather_into copies into caller-owned memory (the library's pinned staging buffer) instead of a
 * new bytes object; when total_bytes exceeds the capacity nothing is copied and the caller
 * retries with a larger buffer. */
static PyObject *gather_impl(PyObject *args, int into)
{
    PyObject *seq_in, *attr;
    unsigned long long address = 0, capacity = 0;
    if (into == 1 ? !PyArg_ParseTuple(args, "OUKK", &seq_in, &attr, &address, &capacity)
                  : !PyArg_ParseTuple(args, "OU", &seq_in, &attr)) return NULL;
    if (into == 2) capacity = 0;                /* lengths only: behaves like gather_into with no room */
    PyObject *seq = PySequence_Fast(seq_in, "expected a sequence of probes");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject **items = PySequence_Fast_ITEMS(seq);
    PyObject *lens_obj = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(n > 0 ? n : 1) * (Py_ssize_t)sizeof(int32_t));
    if (!lens_obj) { Py_DECREF(seq); return NULL; }
    int32_t *lens = (int32_t *)PyBytes_AS_STRING(lens_obj);
    /* pass 1: fetch and check the strings, sum the lengths (references are kept for pass 2) */
    PyObject **strs = (PyObject **)PyMem_Malloc(sizeof(PyObject *) * (size_t)(n > 0 ? n : 1));
    if (!strs) { Py_DECREF(seq); Py_DECREF(lens_obj); return PyErr_NoMemory(); }
    Py_ssize_t got = 0;
    size_t total = 0;
    /* Objects whose class keeps `attr` in a __slots__ member (catch_b200.probe.Probe does): the str pointer is read
     * straight from the member's offset instead of going through the attribute protocol, and the objects a few
     * iterations ahead are prefetched -- a list of 10^6 probes is 10^6 scattered objects plus 10^6 scattered str
     * headers, and the pass is bound by those cache misses.  Any other object takes PyObject_GetAttr. */
    PyTypeObject *fast_type = NULL;
    Py_ssize_t fast_off = 0;
    for (Py_ssize_t i = 0; i < n && i < 4; i++) {
        if (PyUnicode_Check(items[i])) continue;
        PyObject *descr = PyObject_GetAttr((PyObject *)Py_TYPE(items[i]), attr);
        if (!descr) { PyErr_Clear(); break; }
        if (Py_IS_TYPE(descr, &PyMemberDescr_Type)) {
            PyMemberDef *m = ((PyMemberDescrObject *)descr)->d_member;
            if (m->type == Py_T_OBJECT_EX && m->offset > 0) {
                fast_type = Py_TYPE(items[i]);
                fast_off = m->offset;
            }
        }
        Py_DECREF(descr);
        break;
    }
    for (Py_ssize_t i = 0; i < n; i++) {
        if (fast_type) {
            if (i + 16 < n) __builtin_prefetch(items[i + 16]);
            if (i + 8 < n && Py_TYPE(items[i + 8]) == fast_type) {
                PyObject *ahead = *(PyObject **)((char *)items[i + 8] + fast_off);
                if (ahead) __builtin_prefetch(ahead);
            }
        }
        if ((i & 1023) == 1023) {
            /* let a thread that waits for the interpreter lock have it: a helper thread gathering the next
             * grouping must not keep the main thread from issuing its next device call for a whole pass.
             * (The caller keeps the list alive and does not change it meanwhile.) */
            Py_BEGIN_ALLOW_THREADS
            Py_END_ALLOW_THREADS
        }
        PyObject *s = items[i];
        PyObject *direct = NULL;
        if (fast_type && Py_TYPE(s) == fast_type) direct = *(PyObject **)((char *)s + fast_off);
        if (direct && PyUnicode_Check(direct)) {
            s = direct;
            Py_INCREF(s);
        } else if (PyUnicode_Check(s)) {
            Py_INCREF(s);
        } else {
            s = PyObject_GetAttr(s, attr);
            if (!s) goto fail;
            if (!PyUnicode_Check(s)) {
                Py_DECREF(s);
                PyErr_SetString(PyExc_TypeError, "sequence attribute is not a str");
                goto fail;
            }
        }
        strs[got++] = s;
        if (PyUnicode_KIND(s) != PyUnicode_1BYTE_KIND) {
            PyErr_SetString(PyExc_ValueError, "sequences must contain single-byte characters only");
            goto fail;
        }
        const Py_ssize_t len = PyUnicode_GET_LENGTH(s);
        if (len > INT32_MAX) {
            PyErr_SetString(PyExc_ValueError, "sequence too long");
            goto fail;
        }
        lens[i] = (int32_t)len;
        total += (size_t)len;
    }
    {
        /* pass 2: copy into a bytes object of the final size, or into the caller's buffer */
        PyObject *data = into ? PyLong_FromUnsignedLongLong((unsigned long long)total)
                              : PyBytes_FromStringAndSize(NULL, (Py_ssize_t)total);
        if (!data) goto fail;
        char *dst = into ? (char *)(uintptr_t)address : PyBytes_AS_STRING(data);
        const int copy = !into || (into == 1 && total <= (size_t)capacity);
        const void **srcs = NULL;
        if (copy && total >= par_copy_min_bytes() && n >= 2 * PAR_COPY_THREADS)
            srcs = (const void **)malloc(sizeof(void *) * (size_t)n);
        if (srcs) {
            copy_job jobs[PAR_COPY_THREADS];
            pthread_t th[PAR_COPY_THREADS];
            int started[PAR_COPY_THREADS];
            for (Py_ssize_t i = 0; i < n; i++) srcs[i] = PyUnicode_1BYTE_DATA(strs[i]);
            /* contiguous item ranges of about equal byte size */
            Py_ssize_t i = 0;
            size_t off = 0;
            for (int t = 0; t < PAR_COPY_THREADS; t++) {
                const size_t want = total / PAR_COPY_THREADS * (size_t)(t + 1);
                jobs[t].src = srcs; jobs[t].len = lens; jobs[t].dst = dst + off; jobs[t].begin = i;
                while (i < n && (t == PAR_COPY_THREADS - 1 || off + (size_t)lens[i] <= want)) off += (size_t)lens[i++];
                jobs[t].end = i;
            }
            Py_BEGIN_ALLOW_THREADS
            for (int t = 0; t < PAR_COPY_THREADS; t++)
                started[t] = jobs[t].end > jobs[t].begin && pthread_create(&th[t], NULL, copy_worker, &jobs[t]) == 0;
            for (int t = 0; t < PAR_COPY_THREADS; t++) {
                if (started[t]) pthread_join(th[t], NULL);
                else copy_worker(&jobs[t]);             /* empty range, or no thread to be had */
            }
            Py_END_ALLOW_THREADS
            free((void *)srcs);
            for (Py_ssize_t k = 0; k < n; k++) Py_DECREF(strs[k]);
        } else {
            for (Py_ssize_t i = 0; i < n; i++) {
                if (copy && i + 8 < n) __builtin_prefetch(PyUnicode_1BYTE_DATA(strs[i + 8]));
                if (copy) memcpy(dst, PyUnicode_1BYTE_DATA(strs[i]), (size_t)lens[i]);
                dst += lens[i];
                Py_DECREF(strs[i]);
            }
        }
        PyMem_Free(strs);
        Py_DECREF(seq);
        PyObject *ret = PyTuple_Pack(2, data, lens_obj);
        Py_DECREF(data);
        Py_DECREF(lens_obj);
        return ret;
    }
fail:
    for (Py_ssize_t i = 0; i < got; i++) Py_DECREF(strs[i]);
    PyMem_Free(strs);
    Py_DECREF(seq);
    Py_DECREF(lens_obj);
    return NULL;
}

static PyObject *gather(PyObject *self, PyObject *args) { return gather_impl(args, 0); }
static PyObject *gather_into(PyObject *self, PyObject *args) { return gather_impl(args, 1); }
static PyObject *lengths(PyObject *self, PyObject *args) { return gather_impl(args, 2); }


/* ---- FASTA wire format ---------------------------------------------------------------------
 * parse_fasta(data, make_uppercase, replace_degenerate, skip_gaps) -> (list of names, list of sequences)
 * The record loop of the reference's seq_io.read_fasta (catch/utils/seq_io.py:104-175) over the raw bytes of an
 * ASCII file, one pass, GIL released: lines end at '\n', '\r\n' or a lone '\r' (Python's universal newlines);
 * every line is right-stripped of ASCII whitespace (str.rstrip: 0x09-0x0d, 0x1c-0x20); an empty line closes the
 * current record, and the next non-empty line must then be a header (AssertionError otherwise, :141-143); a line
 * that starts with '>' opens a record named by the rest of the line; any other line is appended to the current
 * record after upper-casing, [YRWSMKBDHV] -> N and removal of '-' (:149-154), each switchable like the
 * reference's keyword arguments.  Records come back in file order, repeated names included: the caller's
 * `m[name] = seq` in that order gives the reference's OrderedDict (a repeated header restarts the entry but
 * keeps its first position, :145-146).  The caller guarantees that `data` holds no byte above 0x7f. */
typedef struct { size_t name_b, name_e, seq_b, seq_e; } fasta_rec;

static int is_py_space(unsigned char c) { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); }

/* One span [begin, end) of the file that starts at the beginning of a line (the file start or a header line): its
 * records, transformed sequence bytes written to out[begin ..) (never more than were read). */
typedef struct {
    const unsigned char *p;
    size_t begin, end;
    unsigned char *out;
    const unsigned char *map, *keep;
    fasta_rec *rec;
    size_t n_rec, cap, bad_at;
    int oom;
} fasta_job;

static void *fasta_worker(void *arg)
{
    fasta_job *J = (fasta_job *)arg;
    const unsigned char *p = J->p;
    const size_t n = J->end;
    size_t pos = J->begin, w = J->begin;
    int in_record = 0;
    while (pos < n) {
        /* one line [pos, e), terminator consumed */
        const unsigned char *nl = (const unsigned char *)memchr(p + pos, '\n', n - pos);
        size_t e = nl ? (size_t)(nl - p) : n, next = nl ? e + 1 : n;
        const unsigned char *cr = (const unsigned char *)memchr(p + pos, '\r', e - pos);
        if (cr && !((size_t)(cr - p) + 1 == e && nl)) {      /* a lone '\r' ends the line too */
            e = (size_t)(cr - p);
            next = e + 1;
            if (next < n && p[next] == '\n') next++;
        }
        size_t b = pos;
        pos = next;
        while (e > b && is_py_space(p[e - 1])) e--;
        if (e == b) { in_record = 0; continue; }
        if (!in_record && p[b] != '>') { J->bad_at = b; break; }
        if (p[b] == '>') {
            if (J->n_rec == J->cap) {
                const size_t cap2 = J->cap ? J->cap * 2 : 1024;
                fasta_rec *r2 = (fasta_rec *)realloc(J->rec, cap2 * sizeof(fasta_rec));
                if (!r2) { J->oom = 1; break; }
                J->rec = r2;
                J->cap = cap2;
            }
            if (J->n_rec) J->rec[J->n_rec - 1].seq_e = w;
            fasta_rec *r = &J->rec[J->n_rec++];
            r->name_b = b + 1; r->name_e = e; r->seq_b = w; r->seq_e = w;
            in_record = e - b > 1;      /* an empty name is the reference's "no record" marker (:137-143) */
            continue;
        }
        for (size_t i = b; i < e; i++) {
            const unsigned char c = p[i];
            J->out[w] = J->map[c];
            w += J->keep[c];
        }
    }
    if (J->n_rec) J->rec[J->n_rec - 1].seq_e = w;
    return NULL;
}

#define FASTA_THREADS 6
/* files of at least this many bytes are parsed on several threads (CB_FASTA_PAR_BYTES overrides: the tests set it to 1
 * so that tiny files exercise the cutting) */
static size_t fasta_par_min_bytes(void)
{
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("CB_FASTA_PAR_BYTES");
        long long b = e ? atoll(e) : 0;
        v = b > 0 ? (size_t)b : ((size_t)8 << 20);
    }
    return v;
}

static PyObject *parse_fasta(PyObject *self, PyObject *args)
{
    Py_buffer view;
    int make_upper = 1, replace_degenerate = 1, skip_gaps = 1;
    if (!PyArg_ParseTuple(args, "y*|ppp", &view, &make_upper, &replace_degenerate, &skip_gaps)) return NULL;
    const unsigned char *p = (const unsigned char *)view.buf;
    const size_t n = (size_t)view.len;
    unsigned char map[256], keep[256];
    for (int c = 0; c < 256; c++) {
        unsigned char o = (unsigned char)c;
        if (make_upper && o >= 'a' && o <= 'z') o = (unsigned char)(o - 32);
        if (replace_degenerate && strchr("YRWSMKBDHV", o) && o) o = 'N';
        map[c] = o;
        keep[c] = !(skip_gaps && c == '-');
    }
    unsigned char *out = (unsigned char *)malloc(n ? n : 1);
    if (!out) { PyBuffer_Release(&view); return PyErr_NoMemory(); }
    fasta_job jobs[FASTA_THREADS];
    int n_jobs = 0;
    Py_BEGIN_ALLOW_THREADS
    /* Large files are cut at header lines ('>' right after a line end) into spans that are parsed independently: a
     * header line sets the parser's whole state, and a span's output never outgrows its input. */
    size_t cuts[FASTA_THREADS + 1];
    cuts[0] = 0;
    n_jobs = 1;
    if (n >= fasta_par_min_bytes()) {
        for (int t = 1; t < FASTA_THREADS; t++) {
            size_t i = n / FASTA_THREADS * (size_t)t;
            if (i <= cuts[n_jobs - 1]) i = cuts[n_jobs - 1] + 1;
            for (; i < n; i++) {
                const unsigned char *g = (const unsigned char *)memchr(p + i, '>', n - i);
                if (!g) { i = n; break; }
                i = (size_t)(g - p);
                if (i > 0 && (p[i - 1] == '\n' || p[i - 1] == '\r')) break;
            }
            if (i < n) cuts[n_jobs++] = i;
            else break;
        }
    }
    cuts[n_jobs] = n;
    for (int t = 0; t < n_jobs; t++) {
        jobs[t].p = p; jobs[t].begin = cuts[t]; jobs[t].end = cuts[t + 1]; jobs[t].out = out;
        jobs[t].map = map; jobs[t].keep = keep; jobs[t].rec = NULL; jobs[t].n_rec = 0; jobs[t].cap = 0;
        jobs[t].bad_at = (size_t)-1; jobs[t].oom = 0;
    }
    {
        pthread_t th[FASTA_THREADS];
        int started[FASTA_THREADS];
        for (int t = 1; t < n_jobs; t++) started[t] = pthread_create(&th[t], NULL, fasta_worker, &jobs[t]) == 0;
        fasta_worker(&jobs[0]);
        for (int t = 1; t < n_jobs; t++) {
            if (started[t]) pthread_join(th[t], NULL);
            else fasta_worker(&jobs[t]);
        }
    }
    Py_END_ALLOW_THREADS
    PyObject *names = NULL, *seqs = NULL, *ret = NULL;
    size_t n_rec = 0;
    for (int t = 0; t < n_jobs; t++) {
        if (jobs[t].oom) { PyErr_NoMemory(); goto done; }
        if (jobs[t].bad_at != (size_t)-1) {
            PyErr_Format(PyExc_AssertionError, "FASTA: sequence data without a header at byte %zu", jobs[t].bad_at);
            goto done;
        }
        n_rec += jobs[t].n_rec;
    }
    names = PyList_New((Py_ssize_t)n_rec);
    seqs = PyList_New((Py_ssize_t)n_rec);
    if (!names || !seqs) goto done;
    {
        Py_ssize_t at = 0;
        for (int t = 0; t < n_jobs; t++)
            for (size_t i = 0; i < jobs[t].n_rec; i++) {
                const fasta_rec *r = &jobs[t].rec[i];
                PyObject *nm = PyUnicode_New((Py_ssize_t)(r->name_e - r->name_b), 127);
                PyObject *sq = PyUnicode_New((Py_ssize_t)(r->seq_e - r->seq_b), 127);
                if (!nm || !sq) { Py_XDECREF(nm); Py_XDECREF(sq); goto done; }
                memcpy(PyUnicode_1BYTE_DATA(nm), p + r->name_b, r->name_e - r->name_b);
                memcpy(PyUnicode_1BYTE_DATA(sq), out + r->seq_b, r->seq_e - r->seq_b);
                PyList_SET_ITEM(names, at, nm);
                PyList_SET_ITEM(seqs, at, sq);
                at++;
            }
    }
    ret = PyTuple_Pack(2, names, seqs);
done:
    Py_XDECREF(names);
    Py_XDECREF(seqs);
    free(out);
    for (int t = 0; t < n_jobs; t++) free(jobs[t].rec);
    PyBuffer_Release(&view);
    return ret;
}


/* fasta_stream_block(data, replace_degenerate) -> list
 * One block of an ASCII FASTA file that ends at a line end, for the streaming reader (the reference's
 * seq_io.iterate_fasta, catch/utils/seq_io.py:178-232: no upper-casing, no gap removal, blank lines skipped, header
 * text ignored): the list holds, in file order, None for every header line and a bytes object for every run of
 * sequence lines between two headers (lines right-stripped and concatenated, [YRWSMKBDHV] -> N when asked).  The caller
 * joins the runs of a record across blocks and yields it at the next header. */
static PyObject *fasta_stream_block(PyObject *self, PyObject *args)
{
    Py_buffer view;
    int replace_degenerate = 1;
    if (!PyArg_ParseTuple(args, "y*|p", &view, &replace_degenerate)) return NULL;
    const unsigned char *p = (const unsigned char *)view.buf;
    const size_t n = (size_t)view.len;
    unsigned char map[256];
    for (int c = 0; c < 256; c++) map[c] = (unsigned char)((replace_degenerate && c && strchr("YRWSMKBDHV", c)) ? 'N' : c);
    unsigned char *out = (unsigned char *)malloc(n ? n : 1);
    /* runs: [b, e) in out; a header is recorded as b == (size_t)-1 */
    size_t cap = 256, n_items = 0;
    size_t (*items)[2] = (size_t (*)[2])malloc(cap * sizeof *items);
    if (!out || !items) { free(out); free(items); PyBuffer_Release(&view); return PyErr_NoMemory(); }
    int oom = 0;
    Py_BEGIN_ALLOW_THREADS
    size_t pos = 0, w = 0, run_b = 0;
    while (pos < n && !oom) {
        const unsigned char *nl = (const unsigned char *)memchr(p + pos, '\n', n - pos);
        size_t e = nl ? (size_t)(nl - p) : n, next = nl ? e + 1 : n;
        const unsigned char *cr = (const unsigned char *)memchr(p + pos, '\r', e - pos);
        if (cr && !((size_t)(cr - p) + 1 == e && nl)) {
            e = (size_t)(cr - p);
            next = e + 1;
            if (next < n && p[next] == '\n') next++;
        }
        size_t b = pos;
        pos = next;
        while (e > b && is_py_space(p[e - 1])) e--;
        if (e == b) continue;                                  /* blank line: skipped (:218-220) */
        if (p[b] == '>') {
            if (n_items + 2 > cap) {
                cap *= 2;
                size_t (*i2)[2] = (size_t (*)[2])realloc(items, cap * sizeof *items);
                if (!i2) { oom = 1; break; }
                items = i2;
            }
            if (w > run_b) { items[n_items][0] = run_b; items[n_items][1] = w; n_items++; }
            items[n_items][0] = (size_t)-1; items[n_items][1] = 0; n_items++;
            run_b = w;
            continue;
        }
        for (size_t i = b; i < e; i++) out[w++] = map[p[i]];
    }
    if (!oom && w > run_b) {
        if (n_items + 1 > cap) {
            size_t (*i2)[2] = (size_t (*)[2])realloc(items, (cap + 1) * sizeof *items);
            if (!i2) oom = 1; else items = i2;
        }
        if (!oom) { items[n_items][0] = run_b; items[n_items][1] = w; n_items++; }
    }
    Py_END_ALLOW_THREADS
    PyObject *ret = NULL;
    if (oom) { PyErr_NoMemory(); goto done; }
    ret = PyList_New((Py_ssize_t)n_items);
    if (!ret) goto done;
    for (size_t i = 0; i < n_items; i++) {
        PyObject *it;
        if (items[i][0] == (size_t)-1) { it = Py_None; Py_INCREF(it); }
        else it = PyBytes_FromStringAndSize((const char *)out + items[i][0], (Py_ssize_t)(items[i][1] - items[i][0]));
        if (!it) { Py_CLEAR(ret); goto done; }
        PyList_SET_ITEM(ret, (Py_ssize_t)i, it);
    }
done:
    free(out);
    free(items);
    PyBuffer_Release(&view);
    return ret;
}

/* copy_into(buffer, address) -> int bytes: one contiguous buffer (a ProbeBatch's byte matrix) into caller-owned memory
 * (the library's page-locked staging buffer), on several threads from par_copy_min_bytes() on, GIL released. */
typedef struct { const char *src; char *dst; size_t n; } span_job;
static void *span_worker(void *arg)
{
    span_job *j = (span_job *)arg;
    memcpy(j->dst, j->src, j->n);
    return NULL;
}

static PyObject *copy_into(PyObject *self, PyObject *args)
{
    Py_buffer view;
    unsigned long long address = 0;
    if (!PyArg_ParseTuple(args, "y*K", &view, &address)) return NULL;
    const size_t n = (size_t)view.len;
    char *dst = (char *)(uintptr_t)address;
    const char *src = (const char *)view.buf;
    Py_BEGIN_ALLOW_THREADS
    if (n >= par_copy_min_bytes()) {
        span_job jobs[PAR_COPY_THREADS];
        pthread_t th[PAR_COPY_THREADS];
        int started[PAR_COPY_THREADS];
        const size_t per = (n / PAR_COPY_THREADS + 4095) & ~(size_t)4095;
        for (int t = 0; t < PAR_COPY_THREADS; t++) {
            const size_t b = (size_t)t * per < n ? (size_t)t * per : n;
            const size_t e = b + per < n ? b + per : n;
            jobs[t].src = src + b; jobs[t].dst = dst + b; jobs[t].n = (t == PAR_COPY_THREADS - 1 ? n : e) - b;
            started[t] = jobs[t].n > 0 && pthread_create(&th[t], NULL, span_worker, &jobs[t]) == 0;
        }
        for (int t = 0; t < PAR_COPY_THREADS; t++) {
            if (started[t]) pthread_join(th[t], NULL);
            else if (jobs[t].n > 0) span_worker(&jobs[t]);
        }
    } else {
        memcpy(dst, src, n);
    }
    Py_END_ALLOW_THREADS
    PyBuffer_Release(&view);
    return PyLong_FromSize_t(n);
}

static PyMethodDef methods[] = {
    {"fasta_stream_block", fasta_stream_block, METH_VARARGS,
     "fasta_stream_block(data, replace_degenerate=True) -> [None | bytes, ...] (headers and runs of sequence lines)"},
    {"copy_into", copy_into, METH_VARARGS, "copy_into(buffer, address) -> bytes copied (threaded for large buffers)"},
    {"parse_fasta", parse_fasta, METH_VARARGS,
     "parse_fasta(data, make_uppercase=True, replace_degenerate=True, skip_gaps=True) -> (names, sequences)"},
    {"gather", gather, METH_VARARGS, "gather(seq, attr) -> (bytes data, bytes int32 lengths)"},
    {"lengths", lengths, METH_VARARGS, "lengths(seq, attr) -> (int total_bytes, bytes int32 lengths); nothing is copied"},
    {"gather_into", gather_into, METH_VARARGS,
     "gather_into(seq, attr, address, capacity) -> (int total_bytes, bytes int32 lengths)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", NULL, -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
