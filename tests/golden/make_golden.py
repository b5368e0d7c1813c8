#!/usr/bin/env python3
"""
Regenerates the committed fixtures in tests/golden/ from the read-only reference checkout.
Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs
  fixture_reads.npz        the raw int16 signals + read ids of the reference's 7 single-read
                           fast5 fixtures (tests/fast5_files/*.fast5) and 30 multi-read signals
  reference_goldens.json   the goldens the reference's own tests pin for this path, transcribed
                           from tests/test_classify.py:115-180,:198-296 and
                           tests/test_combine_calls.py:27-51, tests/test_load_fast5s.py:42-72
  oracle_outputs.npz       fp64-oracle outputs on those reads for the three shipped models:
                           per-step softmax rows, merged per-read probabilities and calls
  fast5_fixtures.tar.gz    the reference's 7 single-read fast5 test files + one multi-read file,
                           re-packed (`tar czf` of tests/fast5_files and one file of
                           tests/multi_read_fast5_files) so the fast5 readers can be tested where
                           /root/reference does not exist
  refcode_call_batch.json  outputs of the REFERENCE'S OWN call_batch (imported from
                           /root/reference with h5py/keras/tensorflow stubbed) driven by the
                           oracle's forward pass as `model.predict`
"""
import glob
import json
import os
import pathlib
import sys
import types

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = pathlib.Path('/root/reference')
sys.path.insert(0, str(ROOT))

from deepbinner_b200 import hdf5_lite  # noqa: E402
from oracle import deepbinner_oracle as orc  # noqa: E402

MODELS = ['EXP-NBD103_read_starts', 'EXP-NBD103_read_ends', 'SQK-RBK004_read_starts']


def read_fast5(path):
    with hdf5_lite.open_file(path) as h:
        keys = h.keys()
        if 'Raw' in keys:
            group = h['Raw/Reads'].values()[0]
        else:
            name = [k for k in keys if k.startswith('read_')][0]
            group = h[name + '/Raw']
        return group.attrs['read_id'].decode(), group['Signal'].read()


def import_reference_classify():
    for name in ('h5py', 'keras', 'keras.models', 'tensorflow'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['keras.models'].load_model = None
    sys.modules['keras'].backend = types.SimpleNamespace()
    sys.modules['keras'].models = sys.modules['keras.models']
    sys.path.insert(0, str(REF))
    import deepbinner.classify as ref_classify
    return ref_classify


def main():
    files = sorted(glob.glob(str(REF / 'tests/fast5_files/*.fast5')))
    ids, sigs, names = [], [], []
    for f in files:
        rid, sig = read_fast5(f)
        ids.append(rid)
        sigs.append(sig.astype(np.int16))
        names.append(os.path.basename(f))
    multi = {}
    for f in sorted(glob.glob(str(REF / 'tests/multi_read_fast5_files/*.fast5'))):
        with hdf5_lite.open_file(f) as h:
            for k in sorted(h.keys()):
                multi[k[len('read_'):]] = h[k + '/Raw/Signal'].read().astype(np.int16)
    arrays = {'read_ids': np.array(ids), 'file_names': np.array(names)}
    for i, s in enumerate(sigs):
        arrays['signal_{}'.format(i)] = s
    arrays['multi_ids'] = np.array(sorted(multi))
    for i, k in enumerate(sorted(multi)):
        arrays['multi_signal_{}'.format(i)] = multi[k]
    np.savez_compressed(HERE / 'fixture_reads.npz', **arrays)
    import subprocess
    subprocess.check_call(['tar', 'czf', str(HERE / 'fast5_fixtures.tar.gz'), 'fast5_files',
                           'multi_read_fast5_files/FAK33493_2dac03b8dc7b3757bdcf3b4fed263b60fa5da102_1002.fast5'],
                          cwd=str(REF / 'tests'))

    goldens = {
        'source': 'reference tests/test_classify.py, tests/test_combine_calls.py, tests/test_load_fast5s.py',
        'args': {'batch_size': 128, 'scan_size': 6144, 'score_diff': 0.5},
        'start_only': {  # test_classify.py:115-121
            '63c20e8e-9b10-4ede-9862-9a53eec3c512': '1', '618f68a6-3a9a-45e1-afe0-845172b20349': '1',
            '9bfcf22c-5654-4b4c-b8f7-d3cebd416338': '2', '5ce8d6ab-8c24-43cc-808b-50fb336fda2f': '2',
            '424bfd6b-576c-4e2c-bf86-604c771b5ec9': '3', '177c3867-6812-4476-a6da-9e4d5c43b760': '3',
            '2fbd86a4-029a-45cf-8f18-411d542572ba': '12'},
        'end_only': {    # test_classify.py:134-140
            '63c20e8e-9b10-4ede-9862-9a53eec3c512': '1', '618f68a6-3a9a-45e1-afe0-845172b20349': 'none',
            '9bfcf22c-5654-4b4c-b8f7-d3cebd416338': 'none', '5ce8d6ab-8c24-43cc-808b-50fb336fda2f': '2',
            '424bfd6b-576c-4e2c-bf86-604c771b5ec9': '3', '177c3867-6812-4476-a6da-9e4d5c43b760': '3',
            '2fbd86a4-029a-45cf-8f18-411d542572ba': '12'},
        'both_require_either': {  # test_classify.py:154-160
            '63c20e8e-9b10-4ede-9862-9a53eec3c512': '1', '618f68a6-3a9a-45e1-afe0-845172b20349': '1',
            '9bfcf22c-5654-4b4c-b8f7-d3cebd416338': '2', '5ce8d6ab-8c24-43cc-808b-50fb336fda2f': '2',
            '424bfd6b-576c-4e2c-bf86-604c771b5ec9': '3', '177c3867-6812-4476-a6da-9e4d5c43b760': '3',
            '2fbd86a4-029a-45cf-8f18-411d542572ba': '12'},
        'both_require_both': {    # test_classify.py:174-180
            '63c20e8e-9b10-4ede-9862-9a53eec3c512': '1', '618f68a6-3a9a-45e1-afe0-845172b20349': 'none',
            '9bfcf22c-5654-4b4c-b8f7-d3cebd416338': 'none', '5ce8d6ab-8c24-43cc-808b-50fb336fda2f': '2',
            '424bfd6b-576c-4e2c-bf86-604c771b5ec9': '3', '177c3867-6812-4476-a6da-9e4d5c43b760': '3',
            '2fbd86a4-029a-45cf-8f18-411d542572ba': '12'},
        # test_classify.py:213-217 / :249-253 / :287-296 - the only pinned probabilities (2 d.p.)
        'verbose_row_177c3867': {
            'read_id': '177c3867-6812-4476-a6da-9e4d5c43b760',
            'start': ['0.00', '0.00', '0.00', '1.00'] + ['0.00'] * 9, 'start_call': '3',
            'end': ['0.00', '0.00', '0.00', '1.00'] + ['0.00'] * 9, 'end_call': '3'},
        # test_combine_calls.py:27-51: (start, end) -> either / start / both
        'combine_calls': [
            ['1', '1', '1', '1', '1'], ['none', 'none', 'none', 'none', 'none'],
            ['none', '1', '1', 'none', 'none'], ['1', 'none', '1', '1', 'none'],
            ['1', '2', 'none', 'none', 'none']],
        # test_load_fast5s.py:42-72
        'load_fast5': {
            '177c3867-6812-4476-a6da-9e4d5c43b760': {'len': 4971, 'samples': {'0': 714, '4950': 396}},
            '9bfcf22c-5654-4b4c-b8f7-d3cebd416338': {'len': 4983, 'samples': {'0': 493, '4862': 618}},
            '2fbd86a4-029a-45cf-8f18-411d542572ba': {'len': 5395, 'samples': {'0': 505, '5388': 436}}},
        'n_parameters': 107197,   # test_network_architecture.py:37
    }
    (HERE / 'reference_goldens.json').write_text(json.dumps(goldens, indent=1))

    ref_classify = import_reference_classify()
    out = {}
    refcode = {}
    for m in MODELS:
        w = orc.load_weights(ROOT / 'deepbinner_b200' / 'models' / (m + '.dbnw'), np.float64)
        sides = ['start', 'end'] if 'NBD103' in m else ['start']
        # the reference only ever runs a starts model on 'start' and an ends model on 'end', but
        # both sides of every model are useful regression data
        for side in sides:
            natural = (side == 'start') == ('starts' in m)
            calls, probs, steps = orc.call_batch(w, sigs, side, 6144, 0.5, return_steps=True)
            key = '{}|{}'.format(m, side)
            out[key + '|calls'] = np.array(calls)
            out[key + '|probs'] = np.array(probs, dtype=np.float64)
            out[key + '|steps'] = steps.astype(np.float32)
            if natural:
                args = types.SimpleNamespace(scan_size=6144.0, batch_size=128, score_diff=0.5)
                model = orc.OracleModel(ROOT / 'deepbinner_b200' / 'models' / (m + '.dbnw'))
                rc, rp = ref_classify.call_batch(1024, 13, ids, sigs, model, args, side)
                refcode[key] = {'calls': rc, 'probs': [[float(v) for v in row] for row in rp]}
    np.savez_compressed(HERE / 'oracle_outputs.npz', **out)
    (HERE / 'refcode_call_batch.json').write_text(json.dumps(refcode))
    for k in sorted(refcode):
        print(k, refcode[k]['calls'])


if __name__ == '__main__':
    main()
