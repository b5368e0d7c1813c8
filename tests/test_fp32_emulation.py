"""Host emulation of the fp32 CUDA engine: dbn_fp32_net.cuh compiled with g++ (each kernel phase
looped over all 512 thread ids) must reproduce the oracle.  Checks packing, indexing, padding and
layer semantics of the device code in a container without a GPU."""
import ctypes
import subprocess

import numpy as np
import pytest

from conftest import ROOT, model_path, sliding_windows
from oracle import deepbinner_oracle as orc


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp('emu') / 'emu_fp32.so'
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', str(so),
                           str(ROOT / 'tests' / 'emulation' / 'emu_fp32.cpp')])
    return ctypes.CDLL(str(so))


def test_emulated_network_matches_oracle(emu, fixture_reads):
    _, sigs, _ = fixture_reads
    x = np.concatenate([orc.make_windows(sigs, 1024, s, 'start') for s in (0, 1, 9)]
                       + [sliding_windows(sigs, 10, seed=3)]).astype(np.float32)
    for m in ('EXP-NBD103_read_starts', 'SQK-RBK004_read_starts'):
        blob = open(model_path(m), 'rb').read()
        ref = orc.forward(orc.load_weights(model_path(m)), x)
        out = np.zeros((len(x), 13), np.float32)
        assert emu.emu_predict(blob, ctypes.c_size_t(len(blob)), ctypes.c_void_p(x.ctypes.data),
                               len(x), ctypes.c_void_p(out.ctypes.data)) == 0
        assert np.abs(out - ref).max() < 2e-5


def test_emulated_window_staging_is_bit_exact(emu, fixture_reads):
    _, sigs, _ = fixture_reads
    sigs = sigs[:2] + [sigs[3][:1500], np.full(900, 7, np.int16), np.zeros(0, np.int16),
                       np.array([5], np.int16)]
    row = np.zeros(1024, np.float32)
    for sig in sigs:
        for side in (0, 1):
            r = min(len(sig), 6144 + 512)
            region = np.ascontiguousarray(sig[:r] if side == 0 else sig[len(sig) - r:])
            if region.size == 0:
                region = np.zeros(1, np.int16)
            for s in range(12):
                emu.emu_stage_window(ctypes.c_void_p(region.ctypes.data), r, s, side,
                                     ctypes.c_void_p(row.ctypes.data))
                ref = orc.make_windows([sig], 1024, s, 'start' if side == 0 else 'end')[0]
                assert np.array_equal(row, ref.astype(np.float32))
