#!/usr/bin/env python3
"""Run the split engine once per size under `ncu --metrics gpu__time_duration.sum` to get the front / tail
kernel durations (tools helper; not a benchmark)."""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200.model import B200Model  # noqa: E402

m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
for eng in ('tcgen05', 'tcgen05-split'):
    m.set_engine(eng)
    for n in (296 * 8,):
        x = torch.randn(n, 1024, device='cuda')
        p = torch.zeros(n, 13, device='cuda')
        for _ in range(3):
            m.predict_device(x.data_ptr(), n, p.data_ptr())
        torch.cuda.synchronize()
