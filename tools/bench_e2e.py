#!/usr/bin/env python3
"""End-to-end rate of the fused call_batch entry on packed int16 reads in pinned host memory (the headline
`e2e` of bench.py), for a sweep of pipeline settings.  Usage: python tools/bench_e2e.py [reads]
Settings come from the environment: DEEPBINNER_B200_CALL_CHUNK (windows per pipelined chunk)."""
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import PackedReads  # noqa: E402
from deepbinner_b200.model import B200Model  # noqa: E402


def main():
    shard = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    rng = np.random.RandomState(0)
    reads = (rng.randn(shard, 1024) * 80 + 500).astype(np.int16)
    pinned = torch.from_numpy(reads).pin_memory().numpy()
    m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
    for batch in (4096, 8192, 16384):
        packed = [PackedReads(pinned[a:a + batch]) for a in range(0, shard, batch)]
        for inflight in (2, 3, 4):
            for continuous in (False, True):
                def run(steps):
                    jobs = []
                    for _ in range(steps):
                        for pk in packed:
                            jobs.append(m.call_batch_async(pk, 'start', 512, 0.5))
                            if len(jobs) == inflight:
                                jobs.pop(0).result()
                        if not continuous:
                            while jobs:
                                jobs.pop(0).result()
                    while jobs:
                        jobs.pop(0).result()
                run(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                run(6)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                print('job {:6d} reads, {} in flight, {}: {:.3f} M reads/s'.format(
                    batch, inflight, 'continuous over steps' if continuous else 'drained every step ', 6 * shard / dt / 1e6), flush=True)


if __name__ == '__main__':
    main()
