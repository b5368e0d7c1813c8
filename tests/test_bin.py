"""`deepbinner bin`: binning a FASTQ/FASTA by a classification table, checked
against fixed expectations and - when /root/reference is present - against the reference's own
bin.py run on the same inputs."""
import gzip
import importlib
import pathlib
import sys
import types

import pytest

from deepbinner_b200 import bin as dbin
from deepbinner_b200 import deepbinner as cli

IDS = ['0f4e1b5a-1c1d-4a3b-9c1d-0123456789ab', '1f4e1b5a-1c1d-4a3b-9c1d-0123456789ab',
       '2f4e1b5a-1c1d-4a3b-9c1d-0123456789ab', '3f4e1b5a-1c1d-4a3b-9c1d-0123456789ab']


def make_inputs(tmp_path, fasta=False, gz=False):
    classes = tmp_path / 'classes.tsv'
    classes.write_text('read_ID\tbarcode_call\n{}\t1\n{}\tnone\n{}\t12\n{}\t1\n\nshort\n'.format(*IDS))
    if fasta:
        text = ''.join('>{} runid=x\nACGT{}\n'.format(rid, 'A' * i) for i, rid in enumerate(IDS))
        name = 'reads.fasta'
    else:
        text = ''.join('@{} runid=x\nACGT{}\n+\n!!!!{}\n'.format(rid, 'A' * i, '#' * i)
                       for i, rid in enumerate(IDS))
        name = 'reads.fastq'
    reads = tmp_path / (name + ('.gz' if gz else ''))
    if gz:
        with gzip.open(str(reads), 'wt') as f:
            f.write(text)
    else:
        reads.write_text(text)
    return classes, reads, text


def gunzip(path):
    with gzip.open(str(path), 'rt') as f:
        return f.read()


@pytest.mark.parametrize('fasta,gz', [(False, False), (True, False), (False, True)])
def test_bin_splits_reads_by_barcode(tmp_path, capsys, fasta, gz):
    classes, reads, text = make_inputs(tmp_path, fasta, gz)
    out = tmp_path / 'out'
    cli.main(['bin', '--classes', str(classes), '--reads', str(reads), '--out_dir', str(out)])
    ext = 'fasta' if fasta else 'fastq'
    names = sorted(p.name for p in out.iterdir())
    assert names == ['barcode01.%s.gz' % ext, 'barcode12.%s.gz' % ext, 'unclassified.%s.gz' % ext]
    per = 2 if fasta else 4
    lines = text.splitlines(keepends=True)
    rec = [''.join(lines[i * per:(i + 1) * per]) for i in range(4)]
    assert gunzip(out / ('barcode01.%s.gz' % ext)) == rec[0] + rec[3]
    assert gunzip(out / ('unclassified.%s.gz' % ext)) == rec[1]
    assert gunzip(out / ('barcode12.%s.gz' % ext)) == rec[2]
    stdout = capsys.readouterr().out
    assert '4 total classifications found' in stdout
    assert '  barcode01         2     ' in stdout and '  none              1     ' in stdout


def test_bin_errors(tmp_path):
    classes, reads, _ = make_inputs(tmp_path)
    with pytest.raises(SystemExit) as e:
        dbin.load_classifications(str(tmp_path / 'missing.tsv'))
    assert 'does not exist' in str(e.value)
    bad = tmp_path / 'bad.tsv'
    bad.write_text('{}\tfoo\n'.format(IDS[0]))
    with pytest.raises(SystemExit) as e:
        dbin.load_classifications(str(bad))
    assert 'non-integer bin of foo' in str(e.value)
    out = tmp_path / 'out'
    out.mkdir()
    (out / 'barcode01.fastq.gz').write_text('')
    with pytest.raises(SystemExit) as e:
        cli.main(['bin', '--classes', str(classes), '--reads', str(reads), '--out_dir', str(out)])
    assert 'already exists' in str(e.value)
    noid = tmp_path / 'noid.fastq'
    noid.write_text('@read1\nACGT\n+\n!!!!\n')
    with pytest.raises(SystemExit) as e:
        cli.main(['bin', '--classes', str(classes), '--reads', str(noid), '--out_dir', str(tmp_path / 'o2')])
    assert 'could not find read ID in header' in str(e.value)


def test_bin_reads_missing_from_the_table_are_reported_not_written(tmp_path, capsys):
    classes, reads, text = make_inputs(tmp_path)
    extra = '@9f4e1b5a-1c1d-4a3b-9c1d-0123456789ab\nAC\n+\n!!\n'
    reads.write_text(text + extra)
    out = tmp_path / 'out'
    cli.main(['bin', '--classes', str(classes), '--reads', str(reads), '--out_dir', str(out)])
    assert '  not found         1     ' in capsys.readouterr().out
    assert extra not in ''.join(gunzip(p) for p in out.iterdir())


def test_bin_matches_the_reference_implementation(tmp_path, capsys):
    ref_root = pathlib.Path('/root/reference')
    if not (ref_root / 'deepbinner' / 'bin.py').is_file():
        pytest.skip('reference checkout not present')
    sys.path.insert(0, str(ref_root))
    try:
        for name in [m for m in sys.modules if m == 'deepbinner' or m.startswith('deepbinner.')]:
            del sys.modules[name]
        ref_bin = importlib.import_module('deepbinner.bin')
    except Exception as e:  # noqa: BLE001
        pytest.skip('reference bin.py not importable here: {}'.format(e))
    finally:
        sys.path.remove(str(ref_root))
    classes, reads, _ = make_inputs(tmp_path)
    outs = {}
    for tag, fn in (('ref', ref_bin.bin_reads), ('ours', dbin.bin_reads)):
        out = tmp_path / tag
        fn(types.SimpleNamespace(classes=str(classes), reads=str(reads), out_dir=str(out)))
        text = capsys.readouterr().out.replace(str(out), 'OUT')
        # the reference prints progress at random intervals; keep the stable lines only
        stable = [l for l in text.replace('\r', '\n').splitlines() if not l.startswith('Writing reads')]
        outs[tag] = ({p.name: gunzip(p) for p in out.iterdir()}, stable)
    assert outs['ref'][0] == outs['ours'][0]
    assert outs['ref'][1] == outs['ours'][1]
