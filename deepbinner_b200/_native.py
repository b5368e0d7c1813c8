"""
ctypes binding of libdeepbinner_b200.so (C ABI: include/deepbinner_b200.h), following the reference's
own FFI convention (`dtw_semi_global.py:26-41`: a .so located relative to the module, loaded with
ctypes, C-contiguous numpy arrays, caller-allocated outputs).

Fails loudly (ImportError / RuntimeError) when the library is missing or no B200 is usable - the
product path has no CPU fallback.
"""

import ctypes
import os
import pathlib

import numpy as np

_PKG = pathlib.Path(__file__).resolve().parent
# DEEPBINNER_B200_LIB: load a differently built copy of the same library (kernel A/B experiments)
LIB_PATH = pathlib.Path(os.environ.get('DEEPBINNER_B200_LIB') or _PKG / 'libdeepbinner_b200.so')

DBN_OK = 0
SIDE_START, SIDE_END = 0, 1
ENGINE_FP32, ENGINE_TCGEN05 = 0, 1
ENGINE_NAMES = {ENGINE_FP32: 'fp32', ENGINE_TCGEN05: 'tcgen05'}
ABI_VERSION = 1

_lib = None


class NativeError(RuntimeError):
    pass


def load_library():
    """Load (building first if absent and nvcc exists) the C-ABI library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build
    default_lib = not os.environ.get('DEEPBINNER_B200_LIB')
    stale = default_lib and LIB_PATH.exists() and build.needs_build() and build.have_nvcc()
    if not LIB_PATH.exists() or os.environ.get('DEEPBINNER_B200_REBUILD') or stale:
        # (a library older than csrc/ is rebuilt when nvcc is here; on a box without nvcc it is used as is)
        try:
            build.build_library(force=True)
        except Exception as e:  # noqa: BLE001
            if not LIB_PATH.exists():
                raise ImportError('libdeepbinner_b200.so is missing and could not be built: {}\n'
                                  'deepbinner_b200 has no CPU fallback.'.format(e))
            import warnings
            warnings.warn('libdeepbinner_b200.so is older than its sources and could not be rebuilt: {}'.format(e))
    lib = ctypes.CDLL(str(LIB_PATH))
    c_void_pp = ctypes.POINTER(ctypes.c_void_p)
    sigs = {
        'db_abi_version': (ctypes.c_int, []),
        'db_last_error': (ctypes.c_char_p, []),
        'db_create': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, c_void_pp]),
        'db_destroy': (None, [ctypes.c_void_p]),
        'db_info': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                   ctypes.POINTER(ctypes.c_int)]),
        'db_set_engine': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
        'db_get_engine': (ctypes.c_int, [ctypes.c_void_p]),
        'db_predict_windows': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                              ctypes.c_void_p]),
        'db_predict_windows_f64': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                  ctypes.c_void_p]),
        'db_predict_windows_device': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_int64, ctypes.c_void_p,
                                                     ctypes.c_void_p]),
        'db_call_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p]),
        'db_call_batch_submit': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                ctypes.POINTER(ctypes.c_int)]),
        'db_call_batch_submit_packed': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                       ctypes.POINTER(ctypes.c_int)]),
        'db_call_batch_wait': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
        'db_call_batch_device': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p]),
        'db_zlib_inflate': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
        'db_fast5_read': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
        'db_fast5_list_root': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64,
                                              ctypes.POINTER(ctypes.c_int)]),
        'db_fast5_batch_read': (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int64,
                                               ctypes.POINTER(ctypes.c_void_p)]),
        'db_fast5_batch_read_reads': (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_int64,
                                                     ctypes.POINTER(ctypes.c_void_p)]),
        'db_fast5_batch_read_sides': (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                                     ctypes.POINTER(ctypes.c_void_p)]),
        'db_fast5_batch_rows': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                                               ctypes.POINTER(ctypes.c_void_p)]),
        'db_fast5_batch_get': (ctypes.c_int, [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_void_p)] * 5),
        'db_fast5_batch_free': (None, [ctypes.c_void_p]),
        'db_tc_num_jobs': (ctypes.c_int, [ctypes.c_void_p]),
        'db_tc_job_table': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int]),
        'db_tc_packed': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
        'db_tc_debug_dump': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p]),
        'db_tc_trace': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p]),
        'db_tc_trace_call': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p]),
        'db_last_gpu_ms': (ctypes.c_float, [ctypes.c_void_p]),
        'db_kernel_launches': (ctypes.c_int64, [ctypes.c_void_p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.db_abi_version() != ABI_VERSION:
        raise ImportError('libdeepbinner_b200.so ABI version mismatch')
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ['db_abi_version', 'db_last_error', 'db_create', 'db_destroy', 'db_info',
                    'db_set_engine', 'db_get_engine', 'db_predict_windows',
                    'db_predict_windows_f64', 'db_predict_windows_device', 'db_call_batch',
                    'db_call_batch_submit', 'db_call_batch_submit_packed', 'db_call_batch_wait',
                    'db_call_batch_device', 'db_last_gpu_ms', 'db_kernel_launches',
                    'db_tc_num_jobs', 'db_tc_job_table', 'db_tc_packed', 'db_tc_debug_dump', 'db_tc_trace', 'db_tc_trace_call', 'db_zlib_inflate', 'db_fast5_read',
                    'db_fast5_list_root', 'db_fast5_batch_read', 'db_fast5_batch_read_reads', 'db_fast5_batch_read_sides', 'db_fast5_batch_rows',
                    'db_fast5_batch_get',
                    'db_fast5_batch_free']


def check(rc, what):
    if rc != DBN_OK:
        msg = load_library().db_last_error().decode('utf-8', 'replace')
        raise NativeError('{} failed ({}): {}'.format(what, rc, msg))


def as_ptr(arr):
    return ctypes.c_void_p(arr.ctypes.data)


def require(arr, dtype):
    """C-contiguous array of `dtype` (copying only if needed)."""
    return np.ascontiguousarray(arr, dtype=dtype)
