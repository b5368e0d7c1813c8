#!/usr/bin/env python3
"""Device-resident rate of the fused call_batch kernel (int16 scan regions -> calls) by launch size and stream count."""
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200.model import B200Model  # noqa: E402


def main():
    shard = 65536
    dev = torch.device('cuda')
    m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
    reads = (torch.randn(shard, 1024, device=dev) * 80 + 500).to(torch.int16)
    probs = torch.zeros(shard, m.n_classes, device=dev)
    calls = torch.zeros(shard, dtype=torch.int8, device=dev)
    for batch in (256, 1024, 2048, 8192):
        off = torch.arange(batch + 1, device=dev, dtype=torch.int64) * 1024
        for ns in (1, 3, 4):
            streams = [torch.cuda.Stream() for _ in range(ns)]
            step = torch.zeros((ns, batch, m.n_classes), device=dev)

            def run():
                for b in range(shard // batch):
                    k = b % ns
                    m.call_batch_device(reads.data_ptr() + b * batch * 2048, off.data_ptr(), batch, 'start', 512, 0.5,
                                        probs.data_ptr() + b * batch * m.n_classes * 4, calls.data_ptr() + b * batch,
                                        step[k].data_ptr(), streams[k].cuda_stream)
            run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                run()
            torch.cuda.synchronize()
            print('launches of {:5d} reads on {} stream(s): {:.3f} M reads/s'.format(batch, ns, 4 * shard / (time.perf_counter() - t0) / 1e6), flush=True)


if __name__ == '__main__':
    main()
