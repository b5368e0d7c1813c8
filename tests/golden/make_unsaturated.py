#!/usr/bin/env python3
"""
Harvests UNSATURATED network windows (oracle top-1 probability < 0.99) from the committed fixture
reads (tests/golden/fixture_reads.npz: the reference's 7 single-read + 30 multi-read signals) and
writes tests/golden/unsaturated_windows.npz.  Needs only the repo (no /root/reference):

    python tests/golden/make_unsaturated.py

Why: nearly every real window saturates (top-1 > 0.99), and a saturated softmax row hides operand
precision errors.  SURVEY Appendix C shows the probability error of a reduced-precision engine
lives entirely in the few unsaturated windows, so the parity statistic of the tensor-core engine
(tests/test_gpu_parity.py, bench.py's `parity` block) is taken on a population that is
unsaturated by construction: per model >= 500 windows with top-1 < 0.99, of which >= 100 in
[0.3, 0.7].

A window = z-score (trim_signal.py:61-69) of signal[offset : offset + 1024] of one fixture read -
exactly what call_batch (classify.py:342-357) feeds the network for a full-length slice.  The file
stores (read index, offset) pairs, not samples, plus the fp64 oracle softmax rows:

  reads            names of the arrays in fixture_reads.npz, index = `read` below
  <model>|read     int32 [n]   index into `reads`
  <model>|offset   int32 [n]   first sample of the window
  <model>|probs    float64 [n, 13] softmax rows of oracle/deepbinner_oracle.py:forward (fp64
                   weights, input cast to float32 first as Keras does)
"""
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))

from oracle import deepbinner_oracle as orc  # noqa: E402
from oracle.torch_cpu import TorchCpuModel  # noqa: E402

MODELS = ['EXP-NBD103_read_starts', 'EXP-NBD103_read_ends', 'SQK-RBK004_read_starts']
STRIDE = {'EXP-NBD103_read_ends': 7}   # scan stride in samples per model (default 24): the ends
                                       # model saturates more often, so it is scanned more densely
DEFAULT_STRIDE = 24
WANT = 640            # windows kept per model
WANT_MID = 160        # of which top-1 in [0.3, 0.7] (as many as exist up to this)


def read_names(z):
    names = ['signal_{}'.format(i) for i in range(len(z['read_ids']))]
    names += ['multi_signal_{}'.format(i) for i in range(len(z['multi_ids']))]
    return names


def windows_of(sig, offsets):
    out = np.empty((len(offsets), 1024), dtype=np.float64)
    for i, o in enumerate(offsets):
        out[i] = orc.normalise(sig[o:o + 1024].astype(np.int16))
    return out


def main():
    z = np.load(HERE / 'fixture_reads.npz')
    names = read_names(z)
    out = {'reads': np.array(names)}
    rng = np.random.RandomState(20261017)
    for m in MODELS:
        path = ROOT / 'deepbinner_b200' / 'models' / (m + '.dbnw')
        fast = TorchCpuModel(path)
        cand_read, cand_off, cand_top = [], [], []
        for ri, name in enumerate(names):
            sig = z[name]
            offs = np.arange(0, len(sig) - 1024, STRIDE.get(m, DEFAULT_STRIDE), dtype=np.int64)
            if len(offs) == 0:
                continue
            p = fast.predict(windows_of(sig, offs).astype(np.float32), batch_size=512)
            top = p.max(axis=1)
            sel = np.nonzero(top < 0.985)[0]     # margin: the fp64 oracle decides below
            cand_read += [ri] * len(sel)
            cand_off += list(offs[sel])
            cand_top += list(top[sel])
        cand_read = np.array(cand_read, dtype=np.int32)
        cand_off = np.array(cand_off, dtype=np.int32)
        cand_top = np.array(cand_top)
        mid = np.nonzero((cand_top >= 0.32) & (cand_top <= 0.68))[0]
        rest = np.setdiff1d(np.arange(len(cand_top)), mid)
        take_mid = rng.permutation(mid)[:WANT_MID]
        take_rest = rng.permutation(rest)[:WANT - len(take_mid)]
        take = np.sort(np.concatenate([take_mid, take_rest]))
        w = orc.load_weights(path, np.float64)
        x = np.stack([orc.normalise(z[names[r]][o:o + 1024]) for r, o in zip(cand_read[take], cand_off[take])])
        probs = orc.forward(w, x.astype(np.float32))
        top = probs.max(axis=1)
        keep = top < 0.99
        out[m + '|read'] = cand_read[take][keep]
        out[m + '|offset'] = cand_off[take][keep]
        out[m + '|probs'] = probs[keep]
        n_mid = int(((top[keep] >= 0.3) & (top[keep] <= 0.7)).sum())
        print('{}: candidates {} (mid {}), kept {} unsaturated, {} in [0.3, 0.7]'.format(
            m, len(cand_top), len(mid), int(keep.sum()), n_mid), flush=True)
    np.savez_compressed(HERE / 'unsaturated_windows.npz', **out)


if __name__ == '__main__':
    main()
