#!/usr/bin/env python3
"""Timeline of CTA 0 of the split engine's tail kernel (four windows per CTA): per job and window the
MMA issue window and the epilogue pass window (clock64 stamps of the diagnostics instantiation)."""
import ctypes
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200.model import B200Model  # noqa: E402
from deepbinner_b200 import _native  # noqa: E402

NAMES = ['conv5', 'conv6', 'conv7', 'conv8', 'conv9', 'c12+14', 'conv11', 'conv10f', 'conv15', 'conv13',
         'conv16', 'c17a', 'c17b', 'c17c', 'c17d', 'conv18', 'conv19', 'conv20']


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
    m.set_engine('tcgen05-split')
    x = torch.randn(n, 1024, device='cuda')
    p = torch.zeros(n, 13, device='cuda')
    trace = np.zeros((32, 4, 8), dtype=np.int64)
    for _ in range(3):
        rc = m._lib.db_tc_trace(m._handle, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(p.data_ptr()),
                                _native.as_ptr(trace))
        _native.check(rc, 'db_tc_trace')
    t0 = int(trace[31, 0, 0])
    print('kernel start 0 | inputs landed {}'.format(int(trace[31, 0, 1]) - t0))
    print('job      win | issue_start issue_end | epi_pass_start epi_pass_end | issue_dur epi_dur')
    for j, name in enumerate(NAMES):
        for w in range(4):
            a, b, c, d = [int(v) - t0 if v > 0 else -1 for v in trace[j, w][:4]]
            if a < 0 and c < 0:
                continue
            print('{:8s} {}  | {:8d} {:8d} | {:8d} {:8d} | {:6d} {:6d}'.format(
                name, w, a, b, c, d, b - a if a >= 0 else -1, d - c if c >= 0 else -1))
    print('total', int(trace[:len(NAMES)].max()) - t0)


if __name__ == '__main__':
    main()
