#!/usr/bin/env python3
"""Throughput and profile of the fast5 readers on the committed fixture files:
native C++ batch reader at several thread counts (both sides / start only, own decoder / zlib) vs the
pure-Python reader, and where the time of one file goes (inflate vs everything else).
Usage: python tools/bench_fast5.py"""
import ctypes
import os
import pathlib
import subprocess
import sys
import tarfile
import tempfile
import time
import zlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200 import _native, load_fast5s as lf  # noqa: E402


def inflate_share(files):
    """Time of inflating the signal of every fixture file alone (same deflate level as the files: 1)."""
    lib = _native.load_library()
    out = {}
    for use_zlib in (0, 1):
        total = 0.0
        for f in files:
            _, sig = lf.get_read_id_and_signal(f)
            raw = sig.tobytes()
            comp = np.frombuffer(zlib.compress(raw, 1), dtype=np.uint8)
            dst = np.zeros(len(raw), np.uint8)
            got = ctypes.c_int64()
            t0 = time.perf_counter()
            for _ in range(40):
                lib.db_zlib_inflate(_native.as_ptr(comp), len(comp), _native.as_ptr(dst), len(raw), ctypes.byref(got), use_zlib)
            total += (time.perf_counter() - t0) / 40
        out['zlib' if use_zlib else 'own decoder'] = total / len(files)
    return out


def main():
    if os.environ.get('DEEPBINNER_B200_ZLIB'):
        print('(DEEPBINNER_B200_ZLIB set: the reader inflates with the system zlib)')
    with tempfile.TemporaryDirectory() as d:
        with tarfile.open(ROOT / 'tests' / 'golden' / 'fast5_fixtures.tar.gz') as t:
            t.extractall(d, filter='data')
        singles = sorted(str(p) for p in pathlib.Path(d, 'fast5_files').glob('*.fast5'))
        files = singles * 400
        mb = sum(os.path.getsize(f) for f in files) / 1e6
        samples = sum(len(lf.get_read_id_and_signal(f)[1]) for f in singles) / len(singles)
        print('{} files ({} distinct), {:.1f} KB and {:.0f} samples per file on average, {} host cores'.format(
            len(files), len(singles), 1e3 * mb / len(files), samples, os.cpu_count()))
        t0 = time.perf_counter()
        for f in files[:700]:
            lf.get_read_id_and_signal_python(f)
        py = 700 / (time.perf_counter() - t0)
        print('python reader (1 thread): {:.0f} files/s'.format(py))
        for sides, label in ((3, 'start + end'), (1, 'start only (inflate stops after the scan region)')):
            print('native reader, sides = {} ({}):'.format(sides, label))
            for th in (1, 2, 4, 8, 16, 32):
                if th > 2 * (os.cpu_count() or 1):
                    break
                t0 = time.perf_counter()
                lf.read_fast5_batch_packed(files, keep=6656, threads=th, sides=sides)
                dt = time.perf_counter() - t0
                print('  {:2d} threads: {:6.0f} files/s ({:.0f} MB/s of fast5, {:.1f}x python)'.format(
                    th, len(files) / dt, mb / dt, len(files) / dt / py))
        t0 = time.perf_counter()
        lf.read_fast5_batch_packed(files[:2800], keep=6656, threads=1, sides=3)
        per_file = (time.perf_counter() - t0) / 2800
        share = inflate_share(singles)
        print('one thread, one file: {:.0f} us in total; inflating its signal alone: {}'.format(
            per_file * 1e6, ', '.join('{} {:.0f} us'.format(k, v * 1e6) for k, v in share.items())))
        if not os.environ.get('DEEPBINNER_B200_ZLIB'):
            env = dict(os.environ, DEEPBINNER_B200_ZLIB='1')
            out = subprocess.run([sys.executable, __file__, '--short'], env=env, capture_output=True, text=True).stdout
            print('same reader with the system zlib:')
            print(''.join('  ' + l + '\n' for l in out.splitlines() if 'threads' in l or 'sides' in l), end='')


if __name__ == '__main__':
    main()
