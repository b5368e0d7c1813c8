"""Pins the CPU oracle against every golden the reference's own tests hold for the hot path."""
import json
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, MODELS, REFERENCE, model_path
from oracle import deepbinner_oracle as orc

NATURAL = [('EXP-NBD103_read_starts', 'start'), ('EXP-NBD103_read_ends', 'end'),
           ('SQK-RBK004_read_starts', 'start')]


@pytest.fixture(scope='module')
def oracle_calls(fixture_reads):
    _, sigs, _ = fixture_reads
    out = {}
    for m, side in NATURAL:
        w = orc.load_weights(model_path(m))
        out[m] = orc.call_batch(w, sigs, side, 6144, 0.5)
    return out


def test_parameter_count(reference_goldens):
    # reference tests/test_network_architecture.py:37
    for m in MODELS:
        w = orc.load_weights(model_path(m))
        n = sum(v.size for v in w.values() if isinstance(v, np.ndarray))
        assert n == reference_goldens['n_parameters'] == 107197
        assert w['input_size'] == 1024 and w['n_classes'] == 13


def test_start_and_end_calls_match_reference_tests(fixture_reads, reference_goldens, oracle_calls):
    ids, _, _ = fixture_reads
    start = dict(zip(ids, oracle_calls['EXP-NBD103_read_starts'][0]))
    end = dict(zip(ids, oracle_calls['EXP-NBD103_read_ends'][0]))
    assert start == reference_goldens['start_only']      # test_classify.py:115-121
    assert end == reference_goldens['end_only']          # test_classify.py:134-140
    either = {r: orc.combine_calls(start[r], end[r], require_either=True) for r in ids}
    both = {r: orc.combine_calls(start[r], end[r], require_both=True) for r in ids}
    assert either == reference_goldens['both_require_either']   # :154-160
    assert both == reference_goldens['both_require_both']       # :174-180


def test_verbose_probability_row(fixture_reads, reference_goldens, oracle_calls):
    # test_classify.py:213-217, :249-253, :287-296 - the only pinned probabilities (2 d.p.)
    ids, _, _ = fixture_reads
    g = reference_goldens['verbose_row_177c3867']
    i = ids.index(g['read_id'])
    calls, probs = oracle_calls['EXP-NBD103_read_starts']
    assert ['%.2f' % p for p in probs[i]] == g['start'] and calls[i] == g['start_call']
    calls, probs = oracle_calls['EXP-NBD103_read_ends']
    assert ['%.2f' % p for p in probs[i]] == g['end'] and calls[i] == g['end_call']


def test_combine_calls_truth_table(reference_goldens):
    # tests/test_combine_calls.py:27-51
    for start, end, either, req_start, both in reference_goldens['combine_calls']:
        assert orc.combine_calls(start, end, require_either=True) == either
        assert orc.combine_calls(start, end, require_start=True) == req_start
        assert orc.combine_calls(start, end, require_both=True) == both


def test_committed_oracle_outputs_are_reproducible(fixture_reads, oracle_outputs):
    _, sigs, _ = fixture_reads
    for m, side in NATURAL:
        w = orc.load_weights(model_path(m))
        calls, probs, steps = orc.call_batch(w, sigs, side, 6144, 0.5, return_steps=True)
        key = '{}|{}'.format(m, side)
        assert list(oracle_outputs[key + '|calls']) == calls
        np.testing.assert_allclose(np.array(probs, dtype=np.float64), oracle_outputs[key + '|probs'],
                                   atol=1e-6)
        np.testing.assert_allclose(steps, oracle_outputs[key + '|steps'], atol=1e-6)


def test_survey_appendix_e_vectors(fixture_reads, oracle_outputs):
    # SURVEY Appendix E (fp64 restatement during the survey): read_11206, NBD start model
    ids, _, _ = fixture_reads
    i = ids.index('177c3867-6812-4476-a6da-9e4d5c43b760')
    p = oracle_outputs['EXP-NBD103_read_starts|start|probs'][i]
    assert abs(p[3] - 0.996412) < 2e-6 and abs(p[8] - 2.697e-3) < 2e-6
    steps = oracle_outputs['EXP-NBD103_read_starts|start|steps'][:, i, :]
    assert ''.join('%x' % a for a in steps.argmax(axis=1)) == '330000000000'


def test_outputs_of_reference_call_batch_code(fixture_reads, oracle_calls):
    """tests/golden/refcode_call_batch.json was produced by the REFERENCE'S OWN call_batch
    (classify.py:325-384, imported from /root/reference with h5py/keras/tensorflow stubbed) driving
    the oracle forward pass; the oracle's restated call_batch must agree with it."""
    ref = json.loads((GOLDEN / 'refcode_call_batch.json').read_text())
    for m, side in NATURAL:
        r = ref['{}|{}'.format(m, side)]
        calls, probs = oracle_calls[m]
        assert r['calls'] == calls
        np.testing.assert_allclose(np.array(r['probs']), np.array(probs, dtype=np.float64), atol=2e-6)


@pytest.mark.skipif(not REFERENCE.exists(), reason='reference checkout not present')
def test_live_reference_python_against_oracle(fixture_reads):
    """Run the reference's own windowing / merge / call code live (authoring container only)."""
    for name in ('h5py', 'keras', 'keras.models', 'tensorflow'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['keras.models'].load_model = None
    sys.modules['keras'].backend = types.SimpleNamespace()
    sys.path.insert(0, str(REFERENCE))
    try:
        import deepbinner.classify as ref_classify
        from deepbinner.trim_signal import normalise as ref_normalise
    finally:
        sys.path.remove(str(REFERENCE))
    ids, sigs, _ = fixture_reads
    rng = np.random.RandomState(1)
    for _ in range(20):
        x = rng.randint(-500, 1500, size=rng.randint(0, 50))
        a, b = ref_normalise(x), orc.normalise(x)
        assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
    args = types.SimpleNamespace(scan_size=6144.0, batch_size=128, score_diff=0.5,
                                 require_either=True, require_start=False, require_both=False)
    model = orc.OracleModel(model_path('EXP-NBD103_read_ends'))
    calls, probs = ref_classify.call_batch(1024, 13, ids, sigs, model, args, 'end')
    ocalls, oprobs = orc.call_batch(model.w, sigs, 'end', 6144, 0.5)
    assert calls == ocalls
    np.testing.assert_allclose(np.array(probs, dtype=float), np.array(oprobs, dtype=float), atol=2e-6)
    for p in oprobs:
        assert ref_classify.get_barcode_call_from_probabilities(p, 0.5) == \
            orc.get_barcode_call_from_probabilities(p, 0.5)
        assert np.allclose(ref_classify.make_sum_to_one(list(p)), orc.make_sum_to_one(list(p)))


def test_torch_cpu_baseline_matches_numpy_oracle(fixture_reads):
    from oracle.torch_cpu import TorchCpuModel
    _, sigs, _ = fixture_reads
    x = np.concatenate([orc.make_windows(sigs, 1024, s, 'start') for s in (0, 1, 5)])
    for m in MODELS:
        ref = orc.forward(orc.load_weights(model_path(m)), x.astype(np.float32))
        got = TorchCpuModel(model_path(m)).predict(x)
        assert np.abs(ref - got).max() < 2e-5


def test_tf_edge_semantics_matter():
    """Appendix B: average-pool edge divisor and stride-2 right-only padding are what the oracle
    implements (a regression guard on the restatement itself)."""
    x = np.arange(12, dtype=np.float64).reshape(1, 4, 3)
    y = orc.avg_pool3_same(x)
    assert np.allclose(y[0, 0], (x[0, 0] + x[0, 1]) / 2) and np.allclose(y[0, 3], (x[0, 2] + x[0, 3]) / 2)
    k = np.zeros((3, 1, 1)); k[2, 0, 0] = 1.0      # picks x[2i+2] under (0,1) padding
    z = orc.conv1d_relu(np.arange(1, 9, dtype=np.float64).reshape(1, 8, 1), k, np.zeros(1), stride=2)
    assert z[0, :, 0].tolist() == [3.0, 5.0, 7.0, 0.0]
