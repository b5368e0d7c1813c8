/* CPython helper for the list-of-arrays form of call_batch (reference classify.py:325 takes `signals` as a
 * list of numpy arrays): fills the pointer / length arrays that db_call_batch_submit takes, through the
 * buffer protocol - ~0.1 us per read instead of ~1.5 us of ndarray.ctypes / from_buffer calls in Python, which
 * made batches of 512 reads host-bound.  Built by deepbinner_b200/build.py next to the CUDA library; optional
 * (model.ReadPointers falls back to pure Python when it is missing). */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* fill(sequence, ptrs_address, lens_address) -> number of leading items handled (== len(sequence) when every
 * item is a C-contiguous int16 buffer; the caller converts the rest). */
static PyObject* fill(PyObject* self, PyObject* args) {
    PyObject* seq;
    unsigned long long ptrs_addr, lens_addr;
    if (!PyArg_ParseTuple(args, "OKK", &seq, &ptrs_addr, &lens_addr)) return NULL;
    PyObject* fast = PySequence_Fast(seq, "signals must be a sequence");
    if (!fast) return NULL;
    uint64_t* ptrs = (uint64_t*)(uintptr_t)ptrs_addr;
    int64_t* lens = (int64_t*)(uintptr_t)lens_addr;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    Py_ssize_t i = 0;
    for (; i < n; ++i) {
        Py_buffer view;
        if (PyObject_GetBuffer(PySequence_Fast_GET_ITEM(fast, i), &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
            PyErr_Clear();
            break;
        }
        const int ok = view.itemsize == 2 && view.format && view.format[0] == 'h' && view.format[1] == 0 && view.ndim == 1;
        if (ok) {
            ptrs[i] = (uint64_t)(uintptr_t)view.buf;
            lens[i] = (int64_t)(view.len / 2);
        }
        PyBuffer_Release(&view);
        if (!ok) break;
    }
    Py_DECREF(fast);
    return PyLong_FromSsize_t(i);
}

static PyMethodDef methods[] = {{"fill", fill, METH_VARARGS, "fill(sequence, ptrs_address, lens_address) -> items handled"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fastptr", NULL, -1, methods};
PyMODINIT_FUNC PyInit__fastptr(void) { return PyModule_Create(&module); }
