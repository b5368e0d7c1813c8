"""CPU tests of the callers either side of the hot path: the realtime loop and
the training-data TSV input, driven with stand-in models / loaders (no GPU)."""
import os
import types

import numpy as np
import pytest

from deepbinner_b200 import classify as cls
from deepbinner_b200 import realtime as rt


class FakeModel:
    """Deterministic stand-in with the predict surface: class = (first sample) mod 13."""
    inputs = [types.SimpleNamespace(shape=(None, 1024, 1))]
    outputs = [types.SimpleNamespace(shape=(None, 13))]

    def predict(self, x, batch_size=256):
        x = np.asarray(x).reshape(len(x), -1)
        out = np.full((len(x), 13), 0.001, dtype=np.float32)
        for i, row in enumerate(x):
            k = int(abs(row[:8]).sum() * 1000) % 13
            out[i, k] = 1.0 - 0.012
        return out


def make_args(**kw):
    d = dict(verbose=False, batch_size=4, scan_size=1024, score_diff=0.5, require_either=False,
             require_start=False, require_both=False, stop=True)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_training_data_tsv_input(tmp_path, capsys):
    rng = np.random.RandomState(0)
    path = tmp_path / 'train.tsv'
    with open(path, 'w') as f:
        for i in range(10):
            f.write('{}\t{}\n'.format(i % 3, ','.join(str(v) for v in rng.randint(300, 700, 1024))))
    assert cls.determine_input_type(str(path)) == 'training_data'
    cls.classify_training_data(str(path), FakeModel(), 1024, None, None, 13, make_args(verbose=True))
    out = capsys.readouterr().out.splitlines()
    assert out[0].split('\t')[:3] == ['read_ID', 'barcode_call', 'none']
    rows = out[1:]
    assert len(rows) == 10
    assert rows[0].split('\t')[0] == 'line_1_barcode_0' and rows[9].split('\t')[0] == 'line_10_barcode_0'
    for r in rows:
        cols = r.split('\t')
        assert len(cols) == 2 + 13 and cols[1] in ['none'] + [str(k) for k in range(1, 13)]
    bad = tmp_path / 'bad.txt'
    bad.write_text('hello world\n')
    with pytest.raises(SystemExit):
        cls.determine_input_type(str(bad))
    with pytest.raises(SystemExit):
        cls.determine_input_type(str(tmp_path / 'missing'))


def test_realtime_round_moves_files(tmp_path, monkeypatch, capsys):
    in_dir, out_dir = tmp_path / 'in', tmp_path / 'in' / 'sorted'
    in_dir.mkdir()
    names = ['a.fast5', 'b.fast5', 'c.fast5', 'sub/d.fast5']
    (in_dir / 'sub').mkdir()
    for n in names:
        (in_dir / n).write_bytes(b'x')
    table = {'a.fast5': '1', 'b.fast5': 'none', 'c.fast5': '12', 'd.fast5': '1'}

    def fake_classify(fast5s, *a, **k):
        calls = {os.path.basename(f): table[os.path.basename(f)] for f in fast5s}
        return ({'id_' + n: c for n, c in calls.items()}, {'id_' + os.path.basename(f): f for f in fast5s})

    monkeypatch.setattr(rt, 'classify_fast5_files', fake_classify)
    monkeypatch.setattr(rt, 'load_and_check_models', lambda *a, **k: (object(), 1024, None, None, 13, 1))
    monkeypatch.setattr(rt, 'determine_single_or_multi_fast5s', lambda f: 'single')
    args = make_args(in_dir=str(in_dir), out_dir=str(out_dir), start_model='m', end_model=None)
    rt.realtime(args, poll_seconds=0)
    assert sorted(p.name for p in (out_dir / 'barcode01').iterdir()) == ['a.fast5', 'd.fast5']
    assert [p.name for p in (out_dir / 'unclassified').iterdir()] == ['b.fast5']
    assert [p.name for p in (out_dir / 'barcode12').iterdir()] == ['c.fast5']
    # nested out_dir is excluded from the next scan, so nothing is left to do
    assert rt.look_for_new_fast5s(in_dir, out_dir, True) == []
    assert 'Barcode     Count' in capsys.readouterr().out
    assert rt.get_directory_name('none') == 'unclassified' and rt.get_directory_name('7') == 'barcode07'


def test_realtime_existing_destination_is_ignored(tmp_path, capsys):
    """Reference realtime.py:111-143: a file whose destination exists is reported, put on the ignore
    list and the watcher carries on; only a round in which EVERY move failed for another reason ends it."""
    out = tmp_path / 'out'
    (out / 'barcode03').mkdir(parents=True)
    (out / 'barcode03' / 'x.fast5').write_bytes(b'old')
    src = tmp_path / 'x.fast5'
    src.write_bytes(b'new')
    ignore = set()
    rt.move_classified_fast5s({'r': '3'}, {'r': str(src)}, out, [str(src)], ignore)   # no SystemExit
    assert str(src) in ignore and src.exists()
    assert (out / 'barcode03' / 'x.fast5').read_bytes() == b'old'
    assert 'could not move 1 fast5 file because it already exists' in capsys.readouterr().out
    # a move that fails for another reason (source vanished) is counted; all failed -> exit
    gone = tmp_path / 'gone.fast5'
    with pytest.raises(SystemExit) as e:
        rt.move_classified_fast5s({'g': '1'}, {'g': str(gone)}, out, [str(gone)], ignore)
    assert 'no files were successfully moved' in str(e.value)
    # one of two fails: reported, no exit
    ok = tmp_path / 'ok.fast5'
    ok.write_bytes(b'x')
    rt.move_classified_fast5s({'g': '1', 'k': 'none'}, {'g': str(gone), 'k': str(ok)}, out,
                              [str(gone), str(ok)], ignore)
    assert (out / 'unclassified' / 'ok.fast5').exists()
    assert 'failed to move 1 fast5 file' in capsys.readouterr().out


def test_realtime_multi_read_round_records_reads_and_ignores_the_files(tmp_path, monkeypatch, capsys):
    """Multi-read input (reference realtime.py:89-101): at most 5 files per round, the files stay in
    place and are ignored from then on; the per-read calls go to multi_read_classifications.tsv."""
    in_dir, out_dir = tmp_path / 'in', tmp_path / 'out'
    in_dir.mkdir()
    for i in range(7):
        (in_dir / 'm{}.fast5'.format(i)).write_bytes(b'x')
    rounds = []

    def fake_classify(fast5s, *a, **k):
        rounds.append(list(fast5s))
        assert k['verified_single_read'] is False
        calls, where = {}, {}
        for f in fast5s:
            for r in range(3):
                rid = '{}_read{}'.format(os.path.basename(f), r)
                calls[rid] = str(r + 1) if r else 'none'
                where[rid] = f
        return calls, where

    monkeypatch.setattr(rt, 'classify_fast5_files', fake_classify)
    monkeypatch.setattr(rt, 'load_and_check_models', lambda *a, **k: (object(), 1024, None, None, 13, 1))
    monkeypatch.setattr(rt, 'determine_single_or_multi_fast5s', lambda f: 'multi')
    args = make_args(in_dir=str(in_dir), out_dir=str(out_dir), start_model='m', end_model=None)
    rt.realtime(args, poll_seconds=0)
    assert [len(r) for r in rounds] == [5, 2]                      # 5 files per round, then the rest
    assert len(list(in_dir.glob('*.fast5'))) == 7                  # nothing is moved
    rows = (out_dir / rt.MULTI_TSV).read_text().splitlines()
    assert rows[0] == 'read_ID\tbarcode_call\tfast5_file' and len(rows) == 1 + 7 * 3
    assert rows[1].split('\t')[:2] == ['m0.fast5_read0', 'none']
