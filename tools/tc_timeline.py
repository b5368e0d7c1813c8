#!/usr/bin/env python3
"""Print the per-job timeline of CTA 0 of the tcgen05 kernel (clock64 stamps), run on a B200."""
import ctypes
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200.model import B200Model, tc_num_jobs  # noqa: E402
from deepbinner_b200 import _native  # noqa: E402

NAMES = ['conv2', 'conv3', 'conv4', 'conv5', 'conv6', 'conv7', 'conv8', 'conv9', 'c12+14', 'conv10f',
         'conv15', 'conv11', 'conv13', 'conv16', 'c17a', 'c17b', 'c17c', 'c17d', 'conv18',
         'conv19', 'conv20']


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    call_mode = len(sys.argv) > 2 and sys.argv[2] == 'call'   # fused call_batch form: int16 scan regions, z-score in the prologue
    m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
    m.set_engine('tcgen05')
    x = torch.randn(n, 1024, device='cuda')
    p = torch.zeros(n, 13, device='cuda')
    trace = np.zeros((32, 2, 16), dtype=np.int64)
    samples = (torch.randn(n * 1024, device='cuda') * 80 + 500).to(torch.int16)
    offsets = torch.arange(n + 1, device='cuda', dtype=torch.int64) * 1024
    for _ in range(3):
        if call_mode:
            rc = m._lib.db_tc_trace_call(m._handle, ctypes.c_void_p(samples.data_ptr()), ctypes.c_void_p(offsets.data_ptr()),
                                         n, ctypes.c_void_p(p.data_ptr()), _native.as_ptr(trace))
        else:
            rc = m._lib.db_tc_trace(m._handle, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(p.data_ptr()),
                                    _native.as_ptr(trace))
        _native.check(rc, 'db_tc_trace')
    nj = tc_num_jobs(m)
    misc = trace[31].reshape(-1)
    t0 = misc[0] if misc[0] > 0 else trace[:nj][trace[:nj] > 0].min()
    print('call mode (int16 scan regions, z-score in the prologue)' if call_mode else 'predict mode (fp32 windows)')
    print('kernel start 0 | conv1 w0 done {} | conv1 w1 done {}'.format(int(misc[1] - t0), int(misc[2] - t0)))
    print('job      win | mma_issue_start issue_end | epi_start epi_end | issue_dur epi_dur | epi: params tmem compute fence | mma: wfull0 part0 wfull1 part1')
    for j in range(nj):
        for w in range(2):
            a, b, c, d, e4, e5, e6, e7 = [int(v - t0) if v > 0 else -1 for v in trace[j, w][:8]]
            x8, x9, x10, x11, x12, x13 = [int(v - t0) if v > 0 else -1 for v in trace[j, w][8:14]]
            print('{:8s} {}  | {:8d} {:8d} | {:8d} {:8d} | {:6d} {:6d} | {:5d} {:5d} {:5d} {:5d} {:5d} | {:5d} {:5d} {:5d} {:5d} | top {:6d} waited {:6d} need {:6d}'.format(
                NAMES[j], w, a, b, c, d, b - a, (d - c) if c >= 0 else -1,
                e4 - c, e5 - e4, e6 - e5, e7 - e6, d - e7, x8 - a, x9 - x8, x10 - x9, b - x10, x11, x12, x13))
    print('issuer turn-around (joint jobs): job | weights requested by the loader | top | weights ok | inputs ok | first MMA | last MMA issued | commits done')
    for j in range(8, nj):
        ld = int(trace[j, 1, 15] - t0)
        v = [int(trace[j, 0, k] - t0) if trace[j, 0, k] > 0 else -1 for k in (11, 12, 13, 0, 14, 1)]
        print('  {:8s} | {:7d} | {:7d} | {:7d} | {:7d} | {:7d} | {:7d} | {:7d}'.format(NAMES[j], ld, *v))
    print('total', int(trace[:nj, :, :14].max() - t0))


if __name__ == '__main__':
    main()
