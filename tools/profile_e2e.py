#!/usr/bin/env python3
"""Where does the end-to-end call_batch pipeline lose time?  Captures the GPU timeline of a few steps with
torch.profiler (CUPTI) and prints, per kind of GPU activity, the busy time and the idle gaps of the compute
engine.  Usage: python tools/profile_e2e.py"""
import pathlib
import sys
import time

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import PackedReads  # noqa: E402
from deepbinner_b200.model import B200Model  # noqa: E402


def main():
    shard, batch = 65536, 8192
    rng = np.random.RandomState(0)
    reads = (rng.randn(shard, 1024) * 80 + 500).astype(np.int16)
    pinned = torch.from_numpy(reads).pin_memory().numpy()
    m = B200Model(str(ROOT / 'deepbinner_b200/models/EXP-NBD103_read_starts.dbnw'))
    packed = [PackedReads(pinned[a:a + batch]) for a in range(0, shard, batch)]

    def run(steps):
        jobs = []
        for _ in range(steps):
            for pk in packed:
                jobs.append(m.call_batch_async(pk, 'start', 512, 0.5))
                if len(jobs) == 3:
                    jobs.pop(0).result()
        while jobs:
            jobs.pop(0).result()
    run(2)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        t0 = time.perf_counter()
        run(3)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    if len(sys.argv) > 1:
        prof.export_chrome_trace(sys.argv[1])
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    print('wall {:.2f} ms for {} reads: {:.3f} M reads/s (under the profiler)'.format(wall * 1e3, 3 * shard, 3 * shard / wall / 1e6))
    kinds = {}
    for e in ev:
        k = 'k_tc_forward' if 'k_tc_forward' in e.name else ('k_merge_call' if 'k_merge' in e.name else e.name[:40])
        kinds.setdefault(k, []).append((e.time_range.start, e.time_range.end))
    for k, iv in sorted(kinds.items(), key=lambda kv: -sum(b - a for a, b in kv[1])):
        print('  {:42s} n {:5d} total {:9.2f} ms  mean {:8.1f} us'.format(k, len(iv), sum(b - a for a, b in iv) / 1e3, sum(b - a for a, b in iv) / len(iv)))
    iv = sorted(kinds.get('k_tc_forward', []))
    if iv:
        busy, cur_a, cur_b, gaps = 0, iv[0][0], iv[0][1], []
        for a, b in iv[1:]:
            if a > cur_b:
                busy += cur_b - cur_a
                gaps.append(a - cur_b)
                cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        busy += cur_b - cur_a
        span = iv[-1][1] - iv[0][0]
        print('network kernel: span {:.2f} ms, union busy {:.2f} ms ({:.1f} %), {} gaps, mean gap {:.1f} us, max {:.1f} us'.format(
            span / 1e3, busy / 1e3, 100.0 * busy / span, len(gaps), (sum(gaps) / len(gaps)) if gaps else 0, max(gaps) if gaps else 0))
        # serial sum of kernel durations vs union = how much the launches overlap
        print('sum of kernel durations {:.2f} ms'.format(sum(b - a for a, b in iv) / 1e3))


if __name__ == '__main__':
    main()
