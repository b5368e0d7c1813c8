#!/usr/bin/env python3
"""Host predict pipeline check: one call over many windows (chunk ramp 1024, 2048, ... 8192 on two
streams) must equal the same windows predicted in single-chunk calls; prints a rough host-path rate."""
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200.model import B200Model  # noqa: E402

m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
x = np.random.RandomState(1).randn(40000, 1024).astype(np.float32)
whole = m.predict(x)
parts = np.concatenate([m.predict(x[i:i + 1000]) for i in range(0, len(x), 1000)])
print('max |whole - parts| =', float(np.abs(whole - parts).max()), 'rows sum to 1:', bool(np.allclose(whole.sum(1), 1, atol=1e-5)))
assert np.array_equal(whole, parts)
t0 = time.perf_counter()
m.predict(x)
dt = time.perf_counter() - t0
print('pageable host array: {:.2f} M windows/s'.format(len(x) / dt / 1e6))
