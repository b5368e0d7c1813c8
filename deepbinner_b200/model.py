"""
B200Model: the model object of the classification path.

Stands in for the Keras `Model` that reference `classify.py:86-103 load_trained_model` returns and
that `call_batch` drives at `classify.py:361` (seam b1 of SURVEY section 8b): `.inputs[0].shape`,
`.outputs[0].shape` and `.predict(x, batch_size)`.  It additionally exposes the fused
`call_batch`-level entry (seam b2) that runs windowing + z-score + CNN + merge + call on the GPU.
"""

import ctypes

import numpy as np

from . import _native, weights


class _TensorSpec:
    """Mimics a Keras tensor just enough for `int(model.inputs[0].shape[1])`."""

    def __init__(self, shape):
        self.shape = tuple(shape)


class B200Model:

    def __init__(self, model_file=None, blob=None, device=0, engine=None):
        if blob is None:
            blob = weights.load_blob(model_file)
        self._lib = _native.load_library()
        self._handle = ctypes.c_void_p()
        self._blob = bytes(blob)
        rc = self._lib.db_create(self._blob, len(self._blob), int(device),
                                 ctypes.byref(self._handle))
        _native.check(rc, 'db_create')
        isz, ncl = ctypes.c_int(), ctypes.c_int()
        _native.check(self._lib.db_info(self._handle, ctypes.byref(isz), ctypes.byref(ncl)),
                      'db_info')
        self.input_size = isz.value
        self.n_classes = ncl.value
        self.device = int(device)
        self.model_file = model_file
        self.inputs = [_TensorSpec((None, self.input_size, 1))]
        self.outputs = [_TensorSpec((None, self.n_classes))]
        if engine is not None:
            self.set_engine(engine)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle:
            self._lib.db_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- engine --------------------------------------------------------------------------------
    def set_engine(self, engine):
        if isinstance(engine, str):
            engine = {v: k for k, v in _native.ENGINE_NAMES.items()}[engine]
        _native.check(self._lib.db_set_engine(self._handle, int(engine)), 'db_set_engine')

    @property
    def engine(self):
        return _native.ENGINE_NAMES[self._lib.db_get_engine(self._handle)]

    @property
    def blob(self):
        return self._blob

    # -- seam b1: model.predict (classify.py:361) -----------------------------------------------
    def predict(self, x, batch_size=256, verbose=0):
        """x: [n, input_size, 1] or [n, input_size] float array (the reference passes float64).
        Returns a freshly allocated, writable float32 [n, n_classes] array of softmax rows
        (the caller keeps and mutates rows, classify.py:370-374)."""
        x = np.asarray(x)
        if x.ndim == 3 and x.shape[2] == 1:
            x = x.reshape(x.shape[0], x.shape[1])
        if x.ndim != 2 or x.shape[1] != self.input_size:
            raise ValueError('expected input of shape (n, {}, 1), got {}'
                             .format(self.input_size, x.shape))
        n = x.shape[0]
        probs = np.empty((n, self.n_classes), dtype=np.float32)
        if n == 0:
            return probs
        if x.dtype == np.float64:
            x = _native.require(x, np.float64)
            rc = self._lib.db_predict_windows_f64(self._handle, _native.as_ptr(x), n,
                                                  _native.as_ptr(probs))
        else:
            x = _native.require(x, np.float32)
            rc = self._lib.db_predict_windows(self._handle, _native.as_ptr(x), n,
                                              _native.as_ptr(probs))
        _native.check(rc, 'db_predict_windows')
        return probs

    # -- seam b2: fused call_batch (classify.py:325-384) -------------------------------------------
    def call_batch(self, signals, side, scan_size, score_diff):
        """signals: list of 1-D integer arrays.  Returns (calls int8 [n] with 0 = 'none',
        probabilities float32 [n, n_classes] after make_sum_to_one)."""
        return self.call_batch_async(signals, side, scan_size, score_diff).result()

    def call_batch_async(self, signals, side, scan_size, score_diff):
        """Submit a call_batch job (db_call_batch_submit: the scan regions are gathered into pinned
        staging and all GPU work is enqueued) and return a PendingCalls; `.result()` waits and returns
        what call_batch returns.  `signals` may be freed / reused as soon as this returns.  Up to four
        jobs per model may be pending."""
        side_code = _native.SIDE_START if side == 'start' else _native.SIDE_END
        job = ctypes.c_int(-1)
        rows = None
        if hasattr(signals, 'samples') and hasattr(signals, 'offsets'):
            # already packed by the native fast5 reader (load_fast5s.PackedSignals): whole reads in one
            # buffer; rows of unreadable files are empty reads
            samples, offsets, rows = signals.samples, signals.offsets, signals.rows
            if samples.size == 0:
                samples = np.zeros(1, dtype=np.int16)
            n = len(offsets) - 1
            rc = self._lib.db_call_batch_submit_packed(self._handle, _native.as_ptr(samples), _native.as_ptr(offsets),
                                                       n, side_code, int(scan_size), float(score_diff),
                                                       ctypes.byref(job))
            if len(rows) == n:
                rows = None
        else:
            if not isinstance(signals, ReadPointers):
                signals = ReadPointers(signals)
            n = signals.n
            rc = self._lib.db_call_batch_submit(self._handle, _native.as_ptr(signals.ptrs), _native.as_ptr(signals.lens),
                                                n, side_code, int(scan_size), float(score_diff), ctypes.byref(job))
        _native.check(rc, 'db_call_batch_submit')
        return PendingCalls(self, job.value, n, rows, keep=signals)

    # -- device-resident entry points (used by bench.py for kernel-only timing) --------------------
    def predict_device(self, d_x_ptr, n, d_probs_ptr, stream_ptr=0):
        rc = self._lib.db_predict_windows_device(self._handle, ctypes.c_void_p(d_x_ptr), int(n),
                                                 ctypes.c_void_p(d_probs_ptr),
                                                 ctypes.c_void_p(stream_ptr))
        _native.check(rc, 'db_predict_windows_device')

    def call_batch_device(self, d_samples_ptr, d_offsets_ptr, n_reads, side, scan_size, score_diff,
                          d_probs_ptr, d_calls_ptr, d_step_ptr=0, stream_ptr=0):
        rc = self._lib.db_call_batch_device(
            self._handle, ctypes.c_void_p(d_samples_ptr), ctypes.c_void_p(d_offsets_ptr),
            int(n_reads), _native.SIDE_START if side == 'start' else _native.SIDE_END,
            int(scan_size), float(score_diff), ctypes.c_void_p(d_probs_ptr),
            ctypes.c_void_p(d_calls_ptr), ctypes.c_void_p(d_step_ptr), ctypes.c_void_p(stream_ptr))
        _native.check(rc, 'db_call_batch_device')

    @property
    def last_gpu_ms(self):
        return float(self._lib.db_last_gpu_ms(self._handle))

    @property
    def kernel_launches(self):
        return int(self._lib.db_kernel_launches(self._handle))


try:
    from . import _fastptr          # optional CPython helper (csrc/dbn_fastptr.c, built by build.py)
except ImportError:                 # pure-Python pointer extraction below
    _fastptr = None


def _data_pointer(a):
    """Address of an array's first element.  ctypes' from_buffer / addressof is ~2.5x cheaper than
    ndarray.ctypes.data (0.8 vs 2 us per read: this is the per-read host cost of the list-of-arrays API),
    but needs a writable, non-empty buffer."""
    if a.size and a.flags.writeable:
        return ctypes.addressof(ctypes.c_char.from_buffer(a))
    return a.ctypes.data


class ReadPointers:
    """A list of 1-D integer signals as what db_call_batch_submit takes: an array of int16 pointers and
    an array of lengths (built once per batch, shared by the start and the end model's jobs).  Keeps the
    arrays alive."""

    def __init__(self, signals):
        if _fastptr is not None:
            # C helper (buffer protocol): handles the leading run of C-contiguous int16 arrays - normally all
            arrays = signals if type(signals) is list else list(signals)
            n = len(arrays)
            ptrs, lens = np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.int64)
            done = _fastptr.fill(arrays, ptrs.ctypes.data, lens.ctypes.data) if n else 0
            if done == n:
                self.arrays, self.n, self.ptrs, self.lens = arrays, n, ptrs, lens
                return
        self.arrays = [s if (type(s) is np.ndarray and s.dtype == np.int16 and s.flags.c_contiguous)
                       else np.ascontiguousarray(s, dtype=np.int16) for s in signals]
        self.n = len(self.arrays)
        self.ptrs = np.fromiter(map(_data_pointer, self.arrays), dtype=np.uint64, count=self.n)
        self.lens = np.fromiter((a.size for a in self.arrays), dtype=np.int64, count=self.n)

    def __len__(self):
        return self.n


class PendingCalls:
    """An in-flight call_batch job of a B200Model (db_call_batch_submit .. db_call_batch_wait)."""

    def __init__(self, model, job, n, rows, keep=None):
        self._model, self._job, self._n, self._rows = model, job, n, rows
        self._keep = keep      # page-locked inputs are read by the copy engine until the job completes
        self._out = None

    def result(self):
        if self._out is None:
            m = self._model
            probs = np.empty((self._n, m.n_classes), dtype=np.float32)
            calls = np.empty(self._n, dtype=np.int8)
            rc = m._lib.db_call_batch_wait(m._handle, self._job, _native.as_ptr(probs), _native.as_ptr(calls))
            _native.check(rc, 'db_call_batch_wait')
            if self._rows is not None:
                calls, probs = calls[self._rows], probs[self._rows]
            self._out = (calls, probs)
            self._keep = None
        return self._out


def signals_fit_int16(signals):
    if isinstance(signals, ReadPointers):
        return True
    if hasattr(signals, 'samples') and hasattr(signals, 'offsets'):   # packed by the native reader
        return signals.samples.dtype == np.int16
    for s in signals:
        s = np.asarray(s)
        if s.dtype == np.int16:
            continue
        if s.dtype.kind not in 'iu':
            return False
        if s.size and (s.min() < -32768 or s.max() > 32767):
            return False
    return True


def pack_scan_regions(signals, side, scan_size, input_size):
    """Concatenate the part of each read that call_batch can ever look at - its first ('start') or
    last ('end') scan_size + input_size/2 samples (classify.py:337-349) - as int16 + int64 offsets."""
    region = scan_size + input_size // 2
    parts = []
    offsets = np.zeros(len(signals) + 1, dtype=np.int64)
    for i, s in enumerate(signals):
        s = np.asarray(s)
        piece = s[:region] if side == 'start' else s[max(len(s) - region, 0):]
        parts.append(piece.astype(np.int16, copy=False))
        offsets[i + 1] = offsets[i] + len(piece)
    samples = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int16)
    if samples.size == 0:
        samples = np.zeros(1, dtype=np.int16)
    return np.ascontiguousarray(samples, dtype=np.int16), offsets


def tc_debug_dump(model, x2, job):
    """Diagnostics for tests: run two windows through tcgen05 jobs 0..job and return the raw bytes of
    both shared-memory activation regions (uint8 [2, 98688])."""
    x2 = _native.require(np.asarray(x2).reshape(2, model.input_size), np.float32)
    out = np.zeros((2, 98688), dtype=np.uint8)
    rc = model._lib.db_tc_debug_dump(model._handle, _native.as_ptr(x2), int(job), _native.as_ptr(out))
    _native.check(rc, 'db_tc_debug_dump')
    return out


def tc_num_jobs(model):
    return int(model._lib.db_tc_num_jobs(model._handle))
