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


def have_nvcc():
    import shutil
    cand = _nvcc()
    return bool(os.path.isabs(cand) and os.path.exists(cand) or shutil.which(cand))


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob('*')) + [PKG.parent / 'include' / 'deepbinner_b200.h']
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force=False, verbose=False, defines=(), out=None):
    """`defines` / `out`: build an experimental variant (-D macros) next to the default library."""
    out = pathlib.Path(out) if out else LIB
    if not force and out == LIB and not needs_build():
        return str(LIB)
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false']
    cmd += ['-D' + d for d in defines]
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [str(CSRC / s) for s in SOURCES] + ['-lz', '-o', str(out)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout)
    if verbose:
        print(proc.stdout)
    return str(out)


def build_fastptr(force=False):
    """The optional CPython helper csrc/dbn_fastptr.c -> deepbinner_b200/_fastptr.<abi>.so (gcc; host only)."""
    import sysconfig
    out = PKG / ('_fastptr' + sysconfig.get_config_var('EXT_SUFFIX'))
    src = CSRC / 'dbn_fastptr.c'
    if not force and out.exists() and out.stat().st_mtime >= src.stat().st_mtime:
        return str(out)
    cmd = ['gcc', '-O2', '-shared', '-fPIC', '-I' + sysconfig.get_paths()['include'], str(src), '-o', str(out)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('gcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout)
    return str(out)


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith('--out=')]
    print(build_library(force='--force' in sys.argv or bool(defs), verbose='-v' in sys.argv, defines=defs,
                        out=outs[0] if outs else None))
    if not outs:
        print(build_fastptr(force='--force' in sys.argv))
