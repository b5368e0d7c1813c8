#!/usr/bin/env python3
"""Throughput of the fast5 readers on the committed fixture files: native C++
batch reader at several thread counts vs the pure-Python reader.  Usage: python tools/bench_fast5.py"""
import os
import pathlib
import sys
import tarfile
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200 import load_fast5s as lf  # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as d:
        with tarfile.open(ROOT / 'tests' / 'golden' / 'fast5_fixtures.tar.gz') as t:
            t.extractall(d, filter='data')
        files = sorted(str(p) for p in pathlib.Path(d, 'fast5_files').glob('*.fast5')) * 400
        mb = sum(os.path.getsize(f) for f in files) / 1e6
        t0 = time.perf_counter()
        for f in files[:700]:
            lf.get_read_id_and_signal_python(f)
        py = 700 / (time.perf_counter() - t0)
        print('python reader (1 thread): {:.0f} files/s'.format(py))
        for th in (1, 2, 4, 8, 16, 32):
            if th > 2 * (os.cpu_count() or 1):
                break
            t0 = time.perf_counter()
            lf.read_fast5_batch(files, keep=6656, threads=th)
            dt = time.perf_counter() - t0
            print('native reader, {:2d} threads: {:.0f} files/s ({:.0f} MB/s of fast5, {:.1f}x python)'.format(
                th, len(files) / dt, mb / dt, len(files) / dt / py))


if __name__ == '__main__':
    main()
