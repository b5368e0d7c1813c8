"""CPU tests of the host-side mirror of the reference interface (deepbinner_b200/classify.py etc.)."""
import io
import re
import types

import numpy as np
import pytest

from conftest import GOLDEN, MODELS, REFERENCE, ROOT, model_path
from deepbinner_b200 import classify as cls
from deepbinner_b200 import deepbinner as cli
from deepbinner_b200 import hdf5_lite, load_fast5s, misc, weights
from deepbinner_b200.model import pack_scan_regions, signals_fit_int16
from oracle import deepbinner_oracle as orc


def make_args(**kw):
    d = dict(verbose=False, batch_size=128, scan_size=6144, score_diff=0.5,
             require_either=False, require_start=False, require_both=False)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_combine_calls_truth_table(reference_goldens):
    for start, end, either, req_start, both in reference_goldens['combine_calls']:
        assert cls.combine_calls(start, end, make_args(require_either=True)) == either
        assert cls.combine_calls(start, end, make_args(require_start=True)) == req_start
        assert cls.combine_calls(start, end, make_args(require_both=True)) == both


def test_call_and_renormalise_match_oracle():
    rng = np.random.RandomState(0)
    for _ in range(200):
        p = rng.dirichlet(np.ones(13) * rng.choice([0.05, 0.5, 5]))
        if rng.rand() < 0.2:
            p[3] = p[7]    # ties
        q = cls.make_sum_to_one(list(p))
        assert np.allclose(q, orc.make_sum_to_one(list(p)), rtol=0, atol=1e-15)
        for thr in (0.5, 0.1, 1.0):
            assert cls.get_barcode_call_from_probabilities(q, thr) == \
                orc.get_barcode_call_from_probabilities(q, thr)
    assert cls.get_barcode_call_from_probabilities([0.1, 0.45, 0.45], 0.0001) == 'none'
    assert cls.get_barcode_call_from_probabilities([0.6, 0.4, 0.0], 0.1) == 'none'
    assert cls.get_barcode_call_from_probabilities([0.1, 0.1, 0.8], 0.5) == '2'


def test_check_input_size_messages():
    cls.check_input_size(1024, 6144)
    cls.check_input_size(1024, 512)
    with pytest.raises(SystemExit) as e:     # reference tests/test_classify.py:63-68
        cls.check_input_size(1024, 6143)
    assert '--scan_size must be a multiple' in str(e.value)
    with pytest.raises(SystemExit) as e:
        cls.check_input_size(1023, 6144)
    assert 'must be even' in str(e.value)


def test_output_header_formats(capsys):
    cls.print_output_header(False, True, True, 13)
    cls.print_output_header(True, True, False, 13)
    cls.print_output_header(True, True, True, 13)
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == 'read_ID\tbarcode_call'
    assert lines[1] == 'read_ID\tbarcode_call\tnone\t1\t2\t3\t4\t5\t6\t7\t8\t9\t10\t11\t12'
    assert lines[2] == ('read_ID\tbarcode_call\tstart_none\tstart_1\tstart_2\tstart_3\tstart_4\t'
                        'start_5\tstart_6\tstart_7\tstart_8\tstart_9\tstart_10\tstart_11\tstart_12\t'
                        'start_barcode_call\tend_none\tend_1\tend_2\tend_3\tend_4\tend_5\tend_6\t'
                        'end_7\tend_8\tend_9\tend_10\tend_11\tend_12\tend_barcode_call')


def test_build_windows_matches_oracle(fixture_reads):
    _, sigs, _ = fixture_reads
    sigs = sigs + [np.zeros(0, dtype=np.int16), np.full(700, 5, dtype=np.int16),
                   np.arange(3, dtype=np.int16)]
    for side in ('start', 'end'):
        for s in (0, 1, 7, 11):
            assert np.array_equal(cls.build_windows(sigs, 1024, s, side),
                                  orc.make_windows(sigs, 1024, s, side))


def test_generic_call_batch_with_foreign_model(fixture_reads, reference_goldens):
    """Seam b1: any object with .predict works through the host loop; goldens reproduced."""
    ids, sigs, _ = fixture_reads
    model = orc.OracleModel(model_path('EXP-NBD103_read_starts'))
    calls, probs = cls.call_batch(1024, 13, ids, sigs, model, make_args(), 'start')
    assert dict(zip(ids, calls)) == reference_goldens['start_only']
    ocalls, oprobs = orc.call_batch(model.w, sigs, 'start', 6144, 0.5)
    np.testing.assert_allclose(np.array(probs, dtype=float), np.array(oprobs, dtype=float), atol=2e-6)


def test_predict_output_is_not_aliased_by_generic_call_batch(fixture_reads):
    """The reference keeps row views of the first predict() result and mutates them; our host loop
    must not corrupt a model that returns the same buffer each call."""
    ids, sigs, _ = fixture_reads
    base = orc.OracleModel(model_path('EXP-NBD103_read_starts'))

    class Reusing:
        inputs, outputs = base.inputs, base.outputs
        buf = None

        def predict(self, x, batch_size=256):
            out = base.predict(x, batch_size)
            if self.buf is None:
                self.buf = out
            else:
                self.buf[...] = out
            return self.buf

    calls, _ = cls.call_batch(1024, 13, ids, sigs, Reusing(), make_args(), 'start')
    assert calls == orc.call_batch(base.w, sigs, 'start', 6144, 0.5)[0]


def test_pack_scan_regions(fixture_reads):
    _, sigs, _ = fixture_reads
    for side in ('start', 'end'):
        samples, offsets = pack_scan_regions(sigs, side, 6144, 1024)
        assert samples.dtype == np.int16 and offsets.dtype == np.int64
        for i, s in enumerate(sigs):
            piece = samples[offsets[i]:offsets[i + 1]]
            assert len(piece) == min(len(s), 6656)
            assert np.array_equal(piece, s[:6656] if side == 'start' else s[-len(piece):])
    samples, offsets = pack_scan_regions([np.zeros(0, np.int16)], 'start', 6144, 1024)
    assert offsets.tolist() == [0, 0] and samples.size >= 1
    assert signals_fit_int16([np.array([1, 2, 40000 - 10000])])
    assert not signals_fit_int16([np.array([1, 2, 40000])])
    assert not signals_fit_int16([np.array([1.5])])


def test_summary_table_format():
    out = io.StringIO()
    misc.print_summary_table({'a': '3', 'b': 'none', 'c': '12', 'd': '3'}, output=out)
    assert out.getvalue() == '\nBarcode     Count\n      3         2\n     12         1\n   none         1\n\n'


def test_cli_argument_validation():
    import argparse
    p = argparse.ArgumentParser()
    sub = p.add_subparsers(dest='subparser_name')
    cli.classify_subparser(sub)
    cli.realtime_subparser(sub)
    a = p.parse_args(['classify', '--native', 'x'])
    cli.check_classify_and_realtime_arguments(a)
    assert a.require_either and a.start_model.endswith('EXP-NBD103_read_starts.dbnw')
    assert a.end_model.endswith('EXP-NBD103_read_ends.dbnw')
    assert a.scan_size == 6144 and a.batch_size == 256 and a.score_diff == 0.5
    a = p.parse_args(['realtime', '--in_dir', 'i', '--out_dir', 'o', '--rapid', '--stop'])
    cli.check_classify_and_realtime_arguments(a)
    assert a.end_model is None and a.stop
    for argv, msg in ((['classify', '--native', '--rapid', 'x'], 'only use one model preset'),
                      (['classify', 'x'], 'at least one model'),
                      (['classify', '--native', '-s', 'm', 'x'], 'cannot explicitly specify'),
                      (['classify', '--rapid', '--require_both', 'x'], 'only be used with two models'),
                      (['classify', '--native', '--score_diff', '0', 'x'], '--score_diff must be'),
                      (['classify', '--native', '--require_both', '--require_start', 'x'], 'only one of')):
        with pytest.raises(SystemExit) as e:
            cli.check_classify_and_realtime_arguments(p.parse_args(argv))
        assert msg in str(e.value)


def test_weight_blob_roundtrip_and_validation():
    blob = open(model_path(MODELS[0]), 'rb').read()
    isz, ncl, tensors = weights.unpack_blob(blob)
    assert (isz, ncl) == (1024, 13) and weights.parameter_count(blob) == 107197
    assert weights.pack_blob(isz, ncl, tensors) == blob
    with pytest.raises(weights.ModelFormatError):
        weights.unpack_blob(b'nonsense' * 10)


def test_fast5_reader_goldens(fixture_reads, reference_goldens):
    # reference tests/test_load_fast5s.py:42-72 (lengths and spot samples)
    ids, sigs, _ = fixture_reads
    for rid, g in reference_goldens['load_fast5'].items():
        s = sigs[ids.index(rid)]
        assert len(s) == g['len']
        for pos, val in g['samples'].items():
            assert s[int(pos)] == val


@pytest.mark.skipif(not REFERENCE.exists(), reason='reference checkout not present')
def test_fast5_reader_on_reference_files(fixture_reads, capsys):
    ids, sigs, names = fixture_reads
    d = REFERENCE / 'tests' / 'fast5_files'
    files = load_fast5s.find_all_fast5s(d, verbose=True)
    assert len(files) == 7 and '7 fast5s found' in capsys.readouterr().err
    for rid, sig, name in zip(ids, sigs, names):
        r, s = load_fast5s.get_read_id_and_signal(d / name)
        assert r == rid and s.dtype == np.int16 and np.array_equal(s, sig)
    assert load_fast5s.get_read_id_and_signal(d / 'not_a_real_file.fast5') == (None, None)
    assert load_fast5s.determine_single_or_multi_fast5s(files) == 'single'
    multi = load_fast5s.find_all_fast5s(REFERENCE / 'tests' / 'multi_read_fast5_files')
    assert load_fast5s.determine_single_or_multi_fast5s(multi) == 'multi'
    assert len(load_fast5s.get_root_level_keys(multi[0])) == 10
    assert cls.determine_input_type(str(d)) == 'directory'
    assert cls.determine_input_type(str(d / names[0])) == 'single_fast5'


@pytest.mark.skipif(not REFERENCE.exists(), reason='reference checkout not present')
def test_committed_blobs_equal_reference_model_files():
    for m in MODELS:
        assert weights.load_blob(REFERENCE / 'models' / m) == open(model_path(m), 'rb').read()
    with pytest.raises((weights.ModelFormatError, hdf5_lite.Hdf5Error)):
        weights.load_blob(REFERENCE / 'README.md')


def test_c_abi_library_exports_every_declared_symbol():
    """The .so loads without a GPU and exports exactly what include/deepbinner_b200.h declares."""
    import ctypes
    from deepbinner_b200 import _native, build
    lib_path = build.build_library()
    lib = ctypes.CDLL(lib_path)
    header = (ROOT / 'include' / 'deepbinner_b200.h').read_text()
    declared = re.findall(r'DBN_API\s+[\w\s\*]+?\b(db_\w+)\s*\(', header)
    assert sorted(declared) == sorted(_native.EXPORTED_SYMBOLS) and len(declared) == 32
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.db_abi_version.restype = ctypes.c_int
    assert lib.db_abi_version() == 1


def test_product_never_imports_the_oracle():
    for path in (ROOT / 'deepbinner_b200').rglob('*.py'):
        text = path.read_text()
        assert 'import oracle' not in text and 'from oracle' not in text, path


def test_crafted_weight_blobs_are_rejected_not_read_out_of_bounds():
    """parse_blob (csrc/dbn_weights.h) on hostile input, through a host-only entry of the C ABI: an
    offset / count / dims that would wrap around the bounds checks must give DBN_EFORMAT."""
    import ctypes
    import struct
    from deepbinner_b200 import _native
    lib = _native.load_library()
    good = bytearray(open(model_path(MODELS[0]), 'rb').read())
    out = np.zeros((32, 32), np.int32)
    assert lib.db_tc_job_table(bytes(good), len(good), 0, _native.as_ptr(out), 32) == 21
    entry = struct.Struct('<48sI3IQQ')
    name, ndim, d0, d1, d2, off, count = entry.unpack_from(good, 24)
    for bad in ((name, ndim, d0, d1, d2, 2 ** 64 - 8, count),              # offset + count wraps to a small value
                (name, ndim, d0, d1, d2, off, 2 ** 64 - 1),                 # count alone is absurd
                (name, 3, 2 ** 31, 2 ** 31, 4, off, 0),                     # dims product wraps to 0 == count
                (name, 3, 2 ** 24 + 1, 1, 1, off, 2 ** 24 + 1)):            # implausible dimension
        blob = bytearray(good)
        entry.pack_into(blob, 24, *bad)
        rc = lib.db_tc_job_table(bytes(blob), len(blob), 0, _native.as_ptr(out), 32)
        assert rc == -2, (bad[1:], rc, lib.db_last_error())


def test_read_pointers_with_and_without_the_c_helper(monkeypatch):
    """model.ReadPointers (pointer / length arrays of the list-of-arrays call_batch API): the CPython helper
    (csrc/dbn_fastptr.c, buffer protocol) and the pure-Python fallback give the same arrays; anything that is not
    a C-contiguous int16 array is converted."""
    from deepbinner_b200 import build, model
    build.build_fastptr()
    import importlib
    fast = importlib.import_module('deepbinner_b200._fastptr')
    ro = np.arange(7, dtype=np.int16)
    ro.flags.writeable = False
    signals = [np.arange(100, dtype=np.int16), ro, np.zeros(0, np.int16), np.arange(6, dtype=np.int32), [3, 2, 1],
               np.arange(20, dtype=np.int16)[::2]]
    outs = []
    for helper in (fast, None):
        monkeypatch.setattr(model, '_fastptr', helper)
        r = model.ReadPointers(signals)
        assert r.n == len(signals) and r.lens.tolist() == [100, 7, 0, 6, 3, 10]
        assert all(a.dtype == np.int16 and a.flags.c_contiguous for a in r.arrays)
        assert all(int(p) == a.ctypes.data for p, a in zip(r.ptrs, r.arrays) if a.size)
        outs.append([a.tolist() for a in r.arrays])
        # the fast path keeps the caller's arrays (no copies) when every item qualifies
        plain = [np.arange(5, dtype=np.int16), np.arange(9, dtype=np.int16)]
        r2 = model.ReadPointers(plain)
        assert all(x is y for x, y in zip(r2.arrays, plain)) and r2.lens.tolist() == [5, 9]
    assert outs[0] == outs[1]
