import json
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / 'tests' / 'golden'
MODEL_DIR = ROOT / 'deepbinner_b200' / 'models'
REFERENCE = pathlib.Path('/root/reference')
MODELS = ['EXP-NBD103_read_starts', 'EXP-NBD103_read_ends', 'SQK-RBK004_read_starts']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def _gpu_status():
    """None if a usable B200 is present, else the reason (no library / no device)."""
    try:
        from deepbinner_b200.model import B200Model
        B200Model(str(MODEL_DIR / (MODELS[0] + '.dbnw'))).close()
        return None
    except Exception as e:  # noqa: BLE001
        return str(e)


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not errored) on a box without a usable B200 - unless they were
    asked for explicitly with `-m gpu`, where a missing device must fail loudly."""
    if not any('gpu' in item.keywords for item in items):
        return
    if 'gpu' in (config.getoption('-m') or ''):
        return
    why = _gpu_status()
    if why is None:
        return
    skip = pytest.mark.skip(reason='no usable B200: ' + why.splitlines()[0])
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def model_path(name):
    return str(MODEL_DIR / (name + '.dbnw'))


@pytest.fixture(scope='session')
def fixture_reads():
    """The reference's 7 single-read fast5 fixtures as (read_ids, int16 signals, file names)."""
    z = np.load(GOLDEN / 'fixture_reads.npz')
    ids = [str(x) for x in z['read_ids']]
    sigs = [z['signal_{}'.format(i)] for i in range(len(ids))]
    names = [str(x) for x in z['file_names']]
    return ids, sigs, names


@pytest.fixture(scope='session')
def multi_reads():
    z = np.load(GOLDEN / 'fixture_reads.npz')
    ids = [str(x) for x in z['multi_ids']]
    return ids, [z['multi_signal_{}'.format(i)] for i in range(len(ids))]


@pytest.fixture(scope='session')
def reference_goldens():
    return json.loads((GOLDEN / 'reference_goldens.json').read_text())


@pytest.fixture(scope='session')
def oracle_outputs():
    return np.load(GOLDEN / 'oracle_outputs.npz')


def sliding_windows(signals, count, seed=0, input_size=1024):
    """`count` z-scored windows cut at random offsets from real reads (unsaturated softmaxes live
    here; synthetic gaussian windows are almost all saturated 'none')."""
    from oracle import deepbinner_oracle as orc
    rng = np.random.RandomState(seed)
    out = np.zeros((count, input_size), dtype=np.float64)
    for i in range(count):
        s = signals[rng.randint(len(signals))]
        n = rng.choice([input_size, input_size, input_size, rng.randint(1, input_size)])
        a = rng.randint(0, max(len(s) - n, 1))
        piece = orc.normalise(s[a:a + n])
        if rng.rand() < 0.5:
            out[i, :len(piece)] = piece
        else:
            out[i, input_size - len(piece):] = piece
    return out


def synthetic_signals(n_reads, seed=0, length=1024):
    """Reference's own gaussian random-signal recipe (balance.py:171-174): per read
    mean ~ U(300,600), sd ~ U(10,500), samples = int(N(mean, sd))."""
    rng = np.random.RandomState(seed)
    means = rng.uniform(300, 600, n_reads)
    sds = rng.uniform(10, 500, n_reads)
    lens = [length] * n_reads if np.isscalar(length) else list(length)
    return [np.clip(rng.normal(means[i], sds[i], lens[i]), -32768, 32767).astype(np.int16)
            for i in range(n_reads)]


@pytest.fixture(scope='session')
def fast5_dir(tmp_path_factory):
    """The reference's fast5 test files (7 single-read + 1 multi-read), unpacked from the committed
    archive tests/golden/fast5_fixtures.tar.gz."""
    import tarfile
    d = tmp_path_factory.mktemp('fast5')
    with tarfile.open(GOLDEN / 'fast5_fixtures.tar.gz') as t:
        t.extractall(d, filter='data')
    return d
