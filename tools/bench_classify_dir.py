#!/usr/bin/env python3
"""Directory-level throughput of `classify_fast5_files`: N copies of the fixture
fast5 files, --native preset (start + end models, scan 6144), batch 256.  Reports reads/s including
fast5 parsing, packing, H2D, both networks, merge, calls and TSV formatting.
Usage: python tools/bench_classify_dir.py [copies]"""
import contextlib
import io
import pathlib
import shutil
import sys
import tarfile
import tempfile
import time
import types

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepbinner_b200 import classify as cls  # noqa: E402
from deepbinner_b200 import deepbinner as cli  # noqa: E402


def main():
    copies = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    with tempfile.TemporaryDirectory() as d:
        with tarfile.open(ROOT / 'tests' / 'golden' / 'fast5_fixtures.tar.gz') as t:
            t.extractall(d, filter='data')
        src = sorted(pathlib.Path(d, 'fast5_files').glob('*.fast5'))
        big = pathlib.Path(d, 'big')
        big.mkdir()
        for c in range(copies):
            for f in src:
                shutil.copy(f, big / '{:05d}_{}'.format(c, f.name))
        files = sorted(str(p) for p in big.glob('*.fast5'))
        args = types.SimpleNamespace(batch_size=256, scan_size=6144.0, score_diff=0.5, verbose=False,
                                     require_either=True, require_start=False, require_both=False)
        with contextlib.redirect_stderr(io.StringIO()):
            sm, si, em, ei, out, _ = cls.load_and_check_models(cli.find_native_start_model(),
                                                               cli.find_native_end_model(), 6144)
        for label, n in (('warm-up', 512), ('timed', len(files))):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                cls.classify_fast5_files(files[:n], sm, si, em, ei, out, args, full_output=True,
                                         verified_single_read=True)
            dt = time.perf_counter() - t0
            print('{}: {} fast5 files in {:.2f} s = {:.0f} reads/s ({:.0f} windows/s through two models)'.format(
                label, n, dt, n / dt, 24 * n / dt))


if __name__ == '__main__':
    main()
