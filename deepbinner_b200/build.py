"""
Builds libdeepbinner_b200.so (the C-ABI library, include/deepbinner_b200.h) in-tree with nvcc for
sm_100a.  `python -m deepbinner_b200.build` or `build_library()`; cross-compiles without a GPU.
"""
import os
import pathlib
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
LIB = PKG / 'libdeepbinner_b200.so'
SOURCES = ['dbn_lib.cu', 'dbn_tc.cu', 'dbn_fast5.cpp']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '-shared', '-cudart', 'static']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob('*')) + [PKG.parent / 'include' / 'deepbinner_b200.h']
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return str(LIB)
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false']
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [str(CSRC / s) for s in SOURCES] + ['-lz', '-o', str(LIB)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout)
    if verbose:
        print(proc.stdout)
    return str(LIB)


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
