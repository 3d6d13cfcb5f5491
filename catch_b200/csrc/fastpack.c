/* _fastpack: host-side glue (CPython C API) that gathers the sequences of a list of Probe
 * objects into one contiguous byte buffer plus a length array, in a single pass and without
 * creating intermediate Python objects.  It is the C counterpart of
 *     [p.seq_str for p in probes]; ''.join(...).encode('latin-1'); lengths
 * which dominates the host time of SetCoverFilter.filter() on ~10^5 probes.  No CUDA here; the
 * buffers go to libcatchb200.so (cb_upload_group) through ctypes. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
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
        memcpy(d, j->src[i], (size_t)j->len[i]);
        d += j->len[i];
    }
    return NULL;
}

#define PAR_COPY_MIN_BYTES (8u << 20)
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
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *s = items[i];
        if (PyUnicode_Check(s)) {
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
        if (copy && total >= PAR_COPY_MIN_BYTES && n >= 2 * PAR_COPY_THREADS)
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

static PyMethodDef methods[] = {
    {"gather", gather, METH_VARARGS, "gather(seq, attr) -> (bytes data, bytes int32 lengths)"},
    {"lengths", lengths, METH_VARARGS, "lengths(seq, attr) -> (int total_bytes, bytes int32 lengths); nothing is copied"},
    {"gather_into", gather_into, METH_VARARGS,
     "gather_into(seq, attr, address, capacity) -> (int total_bytes, bytes int32 lengths)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", NULL, -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
