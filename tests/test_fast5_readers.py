"""fast5 readers: the native C++ reader behind the C ABI and its pure-Python twin
against the reference's goldens (tests/test_load_fast5s.py:32-85) and against each other."""
import numpy as np
import pytest

from deepbinner_b200 import load_fast5s as lf


def singles(fast5_dir):
    return sorted(str(p) for p in (fast5_dir / 'fast5_files').glob('*.fast5'))


def test_native_reader_is_loaded():
    assert lf._native_lib() is not None


def test_goldens_of_the_reference_tests(fast5_dir, fixture_reads, reference_goldens, capsys):
    ids, sigs, names = fixture_reads
    files = lf.find_all_fast5s(fast5_dir / 'fast5_files', verbose=True)
    assert len(files) == 7 and '7 fast5s found' in capsys.readouterr().err
    for rid, sig, name in zip(ids, sigs, names):
        for reader in (lf.get_read_id_and_signal, lf.get_read_id_and_signal_python):
            r, s = reader(fast5_dir / 'fast5_files' / name)
            assert r == rid and s.dtype == np.int16 and np.array_equal(s, sig)
    for rid, g in reference_goldens['load_fast5'].items():
        s = sigs[ids.index(rid)]
        assert len(s) == g['len'] and all(s[int(k)] == v for k, v in g['samples'].items())
    for reader in (lf.get_read_id_and_signal, lf.get_read_id_and_signal_python):
        assert reader(fast5_dir / 'fast5_files' / 'not_a_real_file.fast5') == (None, None)
        assert reader(__file__) == (None, None)


def test_single_and_multi_detection(fast5_dir):
    files = singles(fast5_dir)
    multi = [str(p) for p in (fast5_dir / 'multi_read_fast5_files').glob('*.fast5')]
    assert lf.determine_single_or_multi_fast5s(files) == 'single'
    assert lf.determine_single_or_multi_fast5s(multi) == 'multi'
    native, python = lf.get_root_level_keys(multi[0]), lf.get_root_level_keys_python(multi[0])
    assert sorted(native) == sorted(python) and len(native) == 10
    assert all(k.startswith('read_') for k in native)
    assert lf.get_root_level_keys(files[0]) == lf.get_root_level_keys_python(files[0])
    assert lf.get_root_level_keys(__file__) == []
    with pytest.raises(SystemExit):          # load_fast5s.py:36-38
        lf.get_read_id_and_signal(multi[0])


def test_multi_read_files_are_read_natively(fast5_dir, multi_reads):
    """Every read of a multi-read fast5 straight out of the file (native reader and Python twin):
    ids and signals equal the reference fixture's reads (tests/golden/fixture_reads.npz holds the 30
    reads of the reference's three multi-read files), ordered by group name; mixed batches keep file
    order and map every read to its file."""
    ids, sigs = multi_reads
    by_id = dict(zip(ids, sigs))
    multi = sorted(str(p) for p in (fast5_dir / 'multi_read_fast5_files').glob('*.fast5'))
    assert len(multi) == 1
    single = singles(fast5_dir)[0]
    python = lf.get_reads_python(multi[0])
    assert len(python) == 10 and [r for r, _ in python] == sorted(r for r, _ in python)
    for keep in (0, 6144 + 512):
        read_ids, signals, kept = lf.read_fast5_batch_packed([single, multi[0], str(fast5_dir / 'nope.fast5'), multi[0]],
                                                             keep=keep)
        assert len(read_ids) == 21 and kept == [0] + [1] * 10 + [3] * 10
        assert read_ids[1:11] == [r for r, _ in python] == read_ids[11:]
        assert isinstance(signals, lf.PackedSignals) and len(signals.offsets) == 22 + 1   # + the unreadable file's row
        for rid, s, (prid, ps) in zip(read_ids[1:11], signals[1:11], python):
            ref = by_id[rid]
            assert ps.dtype == np.int16 and np.array_equal(ps, ref)
            if keep:
                assert len(s) == min(len(ref), 2 * keep)
                assert np.array_equal(s[:keep], ref[:keep]) and np.array_equal(s[-keep:], ref[-keep:])
            else:
                assert np.array_equal(s, ref)


def test_batch_reader_truncation_keeps_the_scan_regions(fast5_dir, fixture_reads):
    ids, sigs, names = fixture_reads
    files = [str(fast5_dir / 'fast5_files' / n) for n in names] + [str(fast5_dir / 'nope.fast5')]
    keep = 6144 + 512
    for threads in (1, 4):
        got = lf.read_fast5_batch(files, keep=keep, threads=threads)
        assert got[-1] == (None, None)
        for (rid, s), ref_id, ref in zip(got, ids, sigs):
            assert rid == ref_id
            assert np.array_equal(s[:keep], ref[:keep]) and np.array_equal(s[-keep:], ref[-keep:])
            assert len(s) == min(len(ref), 2 * keep)
    full = lf.read_fast5_batch(files[:-1], keep=0)
    assert all(np.array_equal(s, ref) for (_, s), ref in zip(full, sigs))
    assert lf.read_fast5_batch([]) == []


def test_packed_batch_keeps_the_buffer_behind_the_views(fast5_dir, fixture_reads):
    """read_fast5_batch_packed: readable files only, views into one packed buffer, `rows` = their
    positions in the packed arrays (what B200Model.call_batch hands to the C ABI unchanged)."""
    ids, sigs, names = fixture_reads
    files = [str(fast5_dir / 'nope.fast5')] + [str(fast5_dir / 'fast5_files' / n) for n in names]
    keep = 6144 + 512
    read_ids, signals, kept = lf.read_fast5_batch_packed(files, keep=keep)
    assert read_ids == list(ids) and kept == list(range(1, 8))
    assert isinstance(signals, lf.PackedSignals) and signals.rows.tolist() == kept
    assert len(signals.offsets) == len(files) + 1 and signals.offsets[1] == 0   # unreadable: empty row
    for s, row, ref in zip(signals, signals.rows, sigs):
        assert np.array_equal(s, signals.samples[signals.offsets[row]:signals.offsets[row + 1]])
        assert np.array_equal(s[:keep], ref[:keep]) and np.array_equal(s[-keep:], ref[-keep:])


def _inflate(data, capacity, use_zlib=0):
    import ctypes
    from deepbinner_b200 import _native
    lib = _native.load_library()
    src = np.frombuffer(data, dtype=np.uint8)
    if len(src) == 0:
        src = np.zeros(1, np.uint8)
    dst = np.zeros(max(capacity, 1), dtype=np.uint8)
    got = ctypes.c_int64(-1)
    rc = lib.db_zlib_inflate(_native.as_ptr(src), len(data), _native.as_ptr(dst), capacity, ctypes.byref(got), use_zlib)
    return rc, bytes(dst[:max(got.value, 0)])


def test_own_inflate_equals_zlib(fixture_reads, multi_reads):
    """csrc/dbn_inflate.h (the decoder the reader inflates signal chunks with) against zlib: real signals at
    every compression level (stored, fixed and dynamic Huffman blocks), long matches / RLE, incompressible
    and empty inputs, multi-block streams, truncated output buffers; malformed streams are rejected."""
    import zlib
    _, sigs, _ = fixture_reads
    _, msigs = multi_reads
    rng = np.random.RandomState(7)
    payloads = [s.tobytes() for s in sigs[:4] + msigs[:6]]
    payloads += [b'', b'a', b'abc' * 7, bytes(1000), bytes(range(256)) * 300, rng.bytes(70000),
                 (rng.randint(0, 4, 200000).astype(np.uint8)).tobytes(),            # short codes, many matches
                 np.repeat(rng.randint(0, 255, 3000).astype(np.uint8), rng.randint(1, 600, 3000)).tobytes(),   # long runs
                 (rng.normal(500, 40, 150000).astype(np.int16)).tobytes()]
    for data in payloads:
        for level in (0, 1, 6, 9):
            comp = zlib.compress(data, level)
            rc, out = _inflate(comp, len(data) + 16)
            assert rc == 0 and out == data, (len(data), level)
            for cap in {0, 1, len(data) // 3, max(len(data) - 1, 0), len(data)}:      # cut like an HDF5 edge chunk
                rc, out = _inflate(comp, cap)
                assert rc == 0 and out == data[:cap], (len(data), level, cap)
        # several deflate blocks with a sync flush in between (stored empty blocks inside the stream)
        c = zlib.compressobj(6)
        comp = c.compress(data[:len(data) // 2]) + c.flush(zlib.Z_SYNC_FLUSH) + c.compress(data[len(data) // 2:]) + c.flush()
        rc, out = _inflate(comp, len(data) + 16)
        assert rc == 0 and out == data
    # malformed input: never a crash, always an error (or, for damage zlib would only catch by its checksum,
    # the same bytes zlib decodes before failing)
    good = zlib.compress(sigs[0].tobytes(), 1)
    for bad in (good[:2], good[:len(good) // 2], b'\x78\x9c\xff\xff\xff', b'\x00' * 20, good[:10] + bytes(200)):
        rc, _ = _inflate(bad, 100000)
        assert rc == 1
    for _ in range(200):
        corrupt = bytearray(good)
        for k in rng.randint(2, len(good) - 4, 3):
            corrupt[k] ^= 1 << rng.randint(8)
        rc, out = _inflate(bytes(corrupt), len(sigs[0]) * 2 + 64)     # no crash, no out-of-bounds write is the property
        assert rc in (0, 1) and len(out) <= len(sigs[0]) * 2 + 64


def test_start_only_reading_stops_after_the_scan_region(fast5_dir, fixture_reads, multi_reads):
    """sides = 1 (a start-model-only run): reads are cut after `keep` samples - and the decompression stops
    there - with the same leading samples, ids and full lengths as a full read; both layouts."""
    ids, sigs, names = fixture_reads
    mids, msigs = multi_reads
    by_id = dict(zip(list(ids) + list(mids), list(sigs) + list(msigs)))
    files = [str(fast5_dir / 'fast5_files' / n) for n in names]
    files += sorted(str(p) for p in (fast5_dir / 'multi_read_fast5_files').glob('*.fast5'))
    for keep in (512, 6144 + 512):
        read_ids, signals, kept = lf.read_fast5_batch_packed(files, keep=keep, sides=1)
        assert len(read_ids) == 17
        for rid, s in zip(read_ids, signals):
            ref = by_id[rid]
            assert len(s) == min(len(ref), keep) and np.array_equal(s, ref[:keep])
        # end / both sides: the usual [first keep | last keep] form
        for sides in (2, 3):
            _, both, _ = lf.read_fast5_batch_packed(files, keep=keep, sides=sides)
            for rid, s in zip(read_ids, both):
                ref = by_id[rid]
                assert len(s) == min(len(ref), 2 * keep) and np.array_equal(s[-keep:], ref[-keep:])
