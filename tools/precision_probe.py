#!/usr/bin/env python3
"""CPU experiment (oracle-based, test infrastructure): which split-bf16 terms can be dropped on which
layer?  The tensor-core engine computes every conv as A_hi*W_hi + A_hi*W_lo + A_lo*W_hi.  Dropping
A_lo*W_hi on a layer == feeding that layer bf16-rounded activations; dropping A_hi*W_lo == bf16-rounded
weights.  For every layer this prints the max-abs softmax error (vs the fp64 oracle) over the committed
unsaturated windows (tests/golden/unsaturated_windows.npz) when ONLY that layer is degraded, all other
layers exact.  Bar: 1e-3 (north_star); keep-bar used for decisions: 2e-4.

    python tools/precision_probe.py [act|w|both]
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import deepbinner_oracle as orc  # noqa: E402

MODELS = ['EXP-NBD103_read_starts', 'EXP-NBD103_read_ends', 'SQK-RBK004_read_starts']


def bf16_rn(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def forward_with(w, x, degrade_act=(), degrade_w=()):
    """orc.forward with the inputs (degrade_act) / kernels (degrade_w) of the named convs rounded to bf16."""
    real = orc.conv1d_relu
    names = {}

    def conv_hook(t, kernel, bias, stride=1):
        name = names.get(id(kernel))
        if name in degrade_act:
            t = bf16_rn(t)
        if name in degrade_w:
            kernel = bf16_rn(kernel)
        return real(t, kernel, bias, stride)
    for i in range(1, 21):
        names[id(w['conv1d_%d/kernel' % i])] = 'conv1d_%d' % i
    orc.conv1d_relu = conv_hook
    try:
        return orc.forward(w, x)
    finally:
        orc.conv1d_relu = real


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'act'
    u = np.load(ROOT / 'tests/golden/unsaturated_windows.npz')
    z = np.load(ROOT / 'tests/golden/fixture_reads.npz')
    reads = [str(r) for r in u['reads']]
    layers = ['conv1d_%d' % i for i in range(2, 21)]
    groups = {'conv2-4': ['conv1d_2', 'conv1d_3', 'conv1d_4'], 'conv3-4': ['conv1d_3', 'conv1d_4'],
              'conv5-9': ['conv1d_%d' % i for i in range(5, 10)], 'conv10-20': ['conv1d_%d' % i for i in range(10, 21)],
              'all': layers}
    table = {}
    for m in MODELS:
        w = orc.load_weights(ROOT / 'deepbinner_b200/models' / (m + '.dbnw'), np.float64)
        x = np.stack([orc.normalise(z[reads[r]][o:o + 1024]) for r, o in zip(u[m + '|read'], u[m + '|offset'])])
        x = x.astype(np.float32)
        ref = u[m + '|probs']
        for name, sel in [(l, [l]) for l in layers] + list(groups.items()):
            got = forward_with(w, x, degrade_act=sel if what in ('act', 'both') else (),
                               degrade_w=sel if what in ('w', 'both') else ())
            err = np.abs(got - ref).max(axis=1)
            table.setdefault(name, []).append((err.max(), np.percentile(err, 99)))
    print('degraded: {} (bf16-rounded = term dropped); max / p99 softmax error per model'.format(what))
    print('{:10s} {}'.format('layer', '   '.join('{:>22s}'.format(m[-16:]) for m in MODELS)))
    for name, vals in table.items():
        print('{:10s} {}'.format(name, '   '.join('{:10.2e} {:10.2e}'.format(a, b) for a, b in vals)))


if __name__ == '__main__':
    main()
