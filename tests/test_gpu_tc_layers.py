"""Per-layer check of the tcgen05 engine: the split-bf16 activation tensors left in shared memory
after every MMA job are dumped and compared with the oracle's intermediate tensors."""
import numpy as np
import pytest

from conftest import model_path
from oracle import deepbinner_oracle as orc

pytestmark = pytest.mark.gpu


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def decode(region, off, ncg, lp, length, lo_delta):
    """[ncg][lp][8] hi/lo arrays -> float [length, ncg*8]; also returns the two halo rows."""
    def arr(o):
        a = region[o:o + ncg * lp * 16].view(np.uint16).reshape(ncg, lp, 8)
        return bf16_to_f32(a)
    full = arr(off) + arr(off + lo_delta)
    t = full[:, 1:length + 1, :].transpose(1, 0, 2).reshape(length, ncg * 8)
    return t, full[:, 0, :], full[:, length + 1, :]


def decode_parity(act0, w, cg0, ncg):
    """Parity-split concat buffer of both windows in window 0's region: 4 arrays [24][36][8]; window w
    uses rows 18w .. 18w+15, rows 18w+16/17 must be zero."""
    arr_bytes = 24 * 36 * 16
    off = 98688 - 4 * arr_bytes
    def arr(o):
        return bf16_to_f32(act0[o:o + arr_bytes].view(np.uint16).reshape(24, 36, 8))
    ye = arr(off) + arr(off + arr_bytes)
    yo = arr(off + 2 * arr_bytes) + arr(off + 3 * arr_bytes)
    r0 = 18 * w
    y = np.zeros((32, ncg * 8), np.float32)
    y[0::2] = ye[cg0:cg0 + ncg, r0:r0 + 16].transpose(1, 0, 2).reshape(16, ncg * 8)
    y[1::2] = yo[cg0:cg0 + ncg, r0:r0 + 16].transpose(1, 0, 2).reshape(16, ncg * 8)
    return y, ye[:, r0 + 16:r0 + 18, :], None


def decode_stacked(act0, w, lp, length, pitch, lo_delta, halo_row=None):
    """Stacked tail tensors (both windows in window 0's region): [6][lp][8] hi/lo, window w at rows
    1 + pitch*w + i; returns (tensor, halo rows, separator rows)."""
    def arr(o):
        return bf16_to_f32(act0[o:o + 6 * lp * 16].view(np.uint16).reshape(6, lp, 8))
    full = arr(0) + arr(lo_delta)
    r0 = 1 + pitch * w
    t = full[:, r0:r0 + length, :].transpose(1, 0, 2).reshape(length, 48)
    halos = np.stack([full[:, 0, :], full[:, lp - 1 if halo_row is None else halo_row, :]])
    sep = full[:, 1 + length:1 + pitch, :] if (pitch > length and length == 16) else None
    return t, halos, sep


# job index -> (oracle tap, decoder(dump, w)); dump[w] = raw bytes of window w's activation region
JOBS = [
    (0, 'conv2', lambda d, w: decode(d[w], 0, 6, 514, 512, 49344)),
    (1, 'conv3', lambda d, w: decode(d[w], 0, 6, 514, 512, 49344)),
    (2, 'bn2', lambda d, w: decode(d[w], 0, 6, 260, 256, 24960)),      # pooled outputs: pitch L/2 + 4
    (3, 'conv5', lambda d, w: decode(d[w], 0, 2, 258, 256, 8256)),
    (4, 'conv6', lambda d, w: decode(d[w], 0, 6, 258, 256, 24768)),
    (5, 'bn3', lambda d, w: decode(d[w], 0, 6, 132, 128, 12672)),
    (6, 'conv8', lambda d, w: decode(d[w], 0, 6, 130, 128, 12480)),
    (7, 'bn4', lambda d, w: decode(d[w], 0, 6, 68, 64, 6528)),
    (8, 'conv12+14', lambda d, w: decode(d[w], 25728, 4, 66, 64, 4224)),
    (9, 'bn5:0', lambda d, w: decode_parity(d[0], w, 0, 6)),      # average pool folded into conv1d_10
    (10, 'conv15', lambda d, w: decode(d[w], 13056, 6, 66, 64, 6336)),
    (11, 'bn5:48', lambda d, w: decode_parity(d[0], w, 6, 6)),
    (12, 'bn5:96', lambda d, w: decode_parity(d[0], w, 12, 6)),
    (13, 'bn5:144', lambda d, w: decode_parity(d[0], w, 18, 6)),
    (17, 'bn6', lambda d, w: decode_stacked(d[0], w, 36, 16, 18, 3456)),
    (18, 'conv18', lambda d, w: decode_stacked(d[0], w, 36, 16, 18, 3456)),
    (19, 'bn7', lambda d, w: decode_stacked(d[0], w, 21, 8, 9, 2016, halo_row=18)),   # pooled output: two spare rows
]


def test_every_job_against_oracle(fixture_reads, engine='tcgen05'):
    from deepbinner_b200.model import B200Model, tc_debug_dump, tc_num_jobs
    _, sigs, _ = fixture_reads
    name = 'EXP-NBD103_read_starts'
    model = B200Model(model_path(name))
    model.set_engine(engine)      # a tensor-core engine that cannot be selected on a B200 is a failure
    assert tc_num_jobs(model) == 21
    x = orc.make_windows(sigs[2:4], 1024, 1, 'start').astype(np.float32)
    taps = {}
    orc.forward(orc.load_weights(model_path(name)), x, taps=taps)
    failures = []
    for job, tap, dec in JOBS:
        dump = tc_debug_dump(model, x, job)
        for w in range(2):
            got, halo_a, halo_b = dec(dump, w)
            if tap.startswith('bn5:'):
                c0 = int(tap.split(':')[1])
                ref = taps['bn5'][w][:, c0:c0 + 48]
            elif tap == 'conv12+14':     # one MMA job computes both 1x1 bottlenecks
                ref = np.concatenate([taps['conv12'][w], taps['conv14'][w]], axis=1)
            else:
                ref = taps[tap][w]
            scale = np.abs(ref).max() + 1e-30
            err = np.abs(got - ref).max() / scale
            halo = max(np.abs(halo_a).max(), np.abs(halo_b).max() if halo_b is not None else 0.0)
            print('[{}] job {:2d} {:8s} window {}: rel err {:.2e} (max |ref| {:.3g}) halo {:.1e}'.format(
                engine, job, tap, w, err, scale, halo))
            if not (err < 2e-4 and halo == 0.0):
                failures.append((job, tap, w, float(err), float(halo)))
    assert not failures, failures
