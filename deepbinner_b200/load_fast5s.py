"""
fast5 discovery and raw-signal extraction with the semantics of reference `load_fast5s.py`
(`get_read_id_and_signal` :25-49, `find_all_fast5s` :52-64, `determine_single_or_multi_fast5s`
:67-90, `get_root_level_keys` :93-98), on top of the built-in HDF5 subset reader (no h5py here).
"""

import ctypes
import os
import random
import sys

import numpy as np

from . import hdf5_lite


def _native_lib():
    """The C-ABI library's fast5 entry points (csrc/dbn_fast5.cpp), or None if the library cannot be
    loaded (then the pure-Python reader below is used; both give identical results)."""
    try:
        from . import _native
        return _native.load_library()
    except Exception:  # noqa: BLE001
        return None


class PackedSignals(list):
    """A list of per-read int16 signals (views) that also carries the packed buffer they are views
    of: `samples` (all reads of the batch concatenated), `offsets` (int64 [n_packed + 1]) and `rows`
    (for every list entry, its row in the packed arrays - unreadable files are rows without a list
    entry).  `call_batch` hands the packed arrays straight to the C ABI instead of re-packing."""

    def __init__(self, signals, samples, offsets, rows):
        super().__init__(signals)
        self.samples, self.offsets, self.rows = samples, offsets, rows


def read_fast5_batch_packed(fast5_files, keep, threads=None, sides=3):
    """Every read of every readable file, in input order (a multi-read fast5 contributes all its
    reads, ordered by group name): -> (read_ids, signals, kept) with `signals` a PackedSignals and
    `kept[i]` the index of the file read i came from.  `sides`: 1 = only the start of the reads will be
    looked at (reads are cut after `keep` samples and decompression stops there), 2 = end, 3 = both."""
    lib = _native_lib()
    files = [str(f) for f in fast5_files]
    if lib is None:     # pure-Python reader: plain lists
        ids, sigs, kept = [], [], []
        for i, f in enumerate(files):
            for rid, sig in get_reads_python(f):
                if keep > 0 and sides == 1:
                    sig = sig[:keep]
                elif keep > 0 and len(sig) > 2 * keep:
                    sig = np.concatenate([sig[:keep], sig[-keep:]])
                ids.append(rid)
                sigs.append(sig)
                kept.append(i)
        return ids, sigs, kept
    if not files:
        return [], PackedSignals([], np.zeros(0, np.int16), np.zeros(1, np.int64), np.zeros(0, np.int64)), []
    ids, samples, offsets, status, row_file = _native_batch(lib, files, keep, threads, reads=True, sides=sides)
    ok = [r for r in range(len(ids)) if status[r] == 0]
    views = [samples[offsets[r]:offsets[r + 1]] for r in ok]
    return ([ids[r] for r in ok], PackedSignals(views, samples, offsets, np.asarray(ok, dtype=np.int64)),
            [int(row_file[r]) for r in ok])


def _native_batch(lib, files, keep, threads, reads, sides=3):
    """db_fast5_batch_read (one row per file) / db_fast5_batch_read_reads (one row per read) ->
    (read_ids, samples, offsets, status, row_file) with one entry per row."""
    n = len(files)
    arr = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(f) for f in files])
    handle = ctypes.c_void_p()
    threads = threads or min(16, os.cpu_count() or 1)
    if reads and sides != 3:
        rc = lib.db_fast5_batch_read_sides(arr, n, int(threads), int(keep), int(sides), ctypes.byref(handle))
    else:
        entry = lib.db_fast5_batch_read_reads if reads else lib.db_fast5_batch_read
        rc = entry(arr, n, int(threads), int(keep), ctypes.byref(handle))
    if rc != 0:
        raise RuntimeError('db_fast5_batch_read failed')
    try:
        rows, rf = ctypes.c_int64(), ctypes.c_void_p()
        lib.db_fast5_batch_rows(handle, ctypes.byref(rows), ctypes.byref(rf))
        rows = rows.value
        ptrs = [ctypes.c_void_p() for _ in range(5)]
        lib.db_fast5_batch_get(handle, *[ctypes.byref(p) for p in ptrs])
        offsets = np.ctypeslib.as_array(ctypes.cast(ptrs[1], ctypes.POINTER(ctypes.c_int64)), (rows + 1,)).copy()
        total = int(offsets[-1])
        samples = np.ctypeslib.as_array(ctypes.cast(ptrs[0], ctypes.POINTER(ctypes.c_int16)),
                                        (max(total, 1),))[:total].copy()
        ids = ctypes.string_at(ptrs[3], rows * 64)
        status = np.ctypeslib.as_array(ctypes.cast(ptrs[4], ctypes.POINTER(ctypes.c_int32)), (max(rows, 1),))[:rows].copy()
        row_file = np.ctypeslib.as_array(ctypes.cast(rf, ctypes.POINTER(ctypes.c_int32)), (max(rows, 1),))[:rows].copy()
    finally:
        lib.db_fast5_batch_free(handle)
    read_ids = [ids[i * 64:(i + 1) * 64].split(b'\x00')[0].decode() if status[i] == 0 else None
                for i in range(rows)]
    return read_ids, samples, offsets, status, row_file


def read_fast5_batch(fast5_files, keep=0, threads=None):
    """Parse many single-read fast5 files on native host threads.
    -> list of (read_id, int16 signal) or (None, None) per file, in input order.  If keep > 0 only the
    first and last `keep` samples of longer signals are returned (concatenated) - everything
    call_batch can ever look at when keep >= scan_size + input_size/2.  (A multi-read file exits like
    the reference's get_read_id_and_signal; read_fast5_batch_packed returns all reads of such files.)"""
    lib = _native_lib()
    files = [str(f) for f in fast5_files]
    if lib is None:
        out = []
        for f in files:
            rid, sig = get_read_id_and_signal_python(f)
            if sig is not None and keep > 0 and len(sig) > 2 * keep:
                sig = np.concatenate([sig[:keep], sig[-keep:]])
            out.append((rid, sig))
        return out
    if not files:
        return []
    read_ids, samples, offsets, status, _ = _native_batch(lib, files, keep, threads, reads=False)
    if (status == 2).any():
        sys.exit('Error: this entry point reads one-read-per-file fast5s; multi-read files go through '
                 'read_fast5_batch_packed / classify_fast5_files')
    return [(read_ids[i], samples[offsets[i]:offsets[i + 1]]) if status[i] == 0 else (None, None)
            for i in range(len(files))]


def get_read_id_and_signal(fast5_file):
    """-> (read_id str, int16 signal) or (None, None) if the file cannot be read."""
    return read_fast5_batch([fast5_file], threads=1)[0]


def get_read_id_and_signal_python(fast5_file):
    """Pure-Python twin of the native reader (hdf5_lite); kept as the cross-check."""
    try:
        with hdf5_lite.open_file(fast5_file) as h:
            keys = h.keys()
            if 'Raw' in keys:           # old single-read layout: /Raw/Reads/Read_<n>
                group = h['Raw/Reads'].values()[0]
            else:                       # new layout: /read_<uuid>/Raw
                reads = [k for k in keys if k.startswith('read_')]
                if len(reads) > 1:
                    sys.exit('Error: Deepbinner does not (yet) support multi-read fast5 files')
                if not reads:
                    return None, None
                group = h[reads[0] + '/Raw']
            read_id = group.attrs['read_id']
            if isinstance(read_id, bytes):
                read_id = read_id.decode()
            signal = group['Signal'].read()
        return str(read_id), signal
    except (OSError, KeyError, IndexError, ValueError):
        return None, None


def get_reads_python(fast5_file):
    """Pure-Python twin of db_fast5_batch_read_reads for one file: [(read_id, signal), ...] - every
    read of a multi-read file (ordered by group name), one read of a single-read file, [] if unreadable."""
    try:
        with hdf5_lite.open_file(fast5_file) as h:
            keys = h.keys()
            if 'Raw' in keys:
                groups = [h['Raw/Reads'].values()[0]]
            else:
                groups = [h[k + '/Raw'] for k in sorted(k for k in keys if k.startswith('read_'))]
            out = []
            for group in groups:
                read_id = group.attrs['read_id']
                if isinstance(read_id, bytes):
                    read_id = read_id.decode()
                out.append((str(read_id), group['Signal'].read()))
            return out
    except (OSError, KeyError, IndexError, ValueError):
        return []


def find_all_fast5s(directory, verbose=False):
    if verbose:
        print('Looking for fast5 files in {}... '.format(directory), file=sys.stderr, end='',
              flush=True)
    found = [os.path.join(root, name)
             for root, _, names in os.walk(str(directory)) for name in names
             if name.endswith('.fast5')]
    if verbose:
        print('{} {} found'.format(len(found), 'fast5' if len(found) == 1 else 'fast5s'),
              file=sys.stderr)
    return found


def get_root_level_keys(fast5_file):
    lib = _native_lib()
    if lib is not None:
        buf = ctypes.create_string_buffer(1 << 20)
        count = ctypes.c_int()
        if lib.db_fast5_list_root(os.fsencode(str(fast5_file)), buf, len(buf), ctypes.byref(count)) != 0:
            return []
        return [x.decode() for x in buf.raw.split(b'\x00')[:count.value]]
    return get_root_level_keys_python(fast5_file)


def get_root_level_keys_python(fast5_file):
    try:
        with hdf5_lite.open_file(fast5_file) as h:
            return h.keys()
    except (OSError, KeyError, IndexError, ValueError):
        return []


def determine_single_or_multi_fast5s(fast5s):
    """Inspect up to five randomly chosen files -> 'single' or 'multi' (exits on an old/multi mix)."""
    sample = list(fast5s)
    random.shuffle(sample)
    kinds = set()
    for path in sample[:5]:
        keys = get_root_level_keys(path)
        if 'Raw' in keys:
            kinds.add('single-old')
            continue
        n_reads = sum(1 for k in keys if k.startswith('read_'))
        if n_reads == 1:
            kinds.add('single-new')
        elif n_reads > 1:
            kinds.add('multi')
    if 'multi' in kinds and 'single-old' in kinds:
        sys.exit('Error: your reads appear to be a mixture of old and new formats. Deepbinner '
                 'can handle one or the other, but not both at once.')
    return 'multi' if 'multi' in kinds else 'single'
