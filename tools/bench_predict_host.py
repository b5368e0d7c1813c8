#!/usr/bin/env python3
"""Rate of the host-array predict seam (B200Model.predict, reference classify.py:361) for pageable numpy arrays:
n = 256 float64 windows per call (what the reference's call_batch passes) and one large float32 / float64 array.
DEEPBINNER_B200_NO_STAGE=1 lets the driver stage the pageable array instead of the library's thread pool."""
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200.model import B200Model  # noqa: E402


def rate(f, n, reps):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return reps * n / (time.perf_counter() - t0)


def main():
    m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
    rng = np.random.RandomState(0)
    x32 = rng.randn(65536, 1024).astype(np.float32)
    x64 = x32[:16384].astype(np.float64)
    a = np.ascontiguousarray(x64[:256, :, None])
    print('{}: n=256 float64 per call {:.3f} M/s | 16384 float64 {:.3f} M/s | 65536 float32 pageable {:.3f} M/s'.format(
        'driver staging' if os.environ.get('DEEPBINNER_B200_NO_STAGE') else 'pool staging  ',
        rate(lambda: m.predict(a, batch_size=256), 256, 300) / 1e6,
        rate(lambda: m.predict(x64), 16384, 5) / 1e6,
        rate(lambda: m.predict(x32), 65536, 5) / 1e6), flush=True)


if __name__ == '__main__':
    main()
