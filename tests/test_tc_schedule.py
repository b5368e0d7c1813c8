"""Host-side checks of the tcgen05 engine's MMA job table (csrc/dbn_tc.cu build_jobs), dumped through
the C ABI without a GPU (db_tc_job_table).

The joint phase lets the MMA issuer run ahead of the epilogue warps: every joint job carries `need`,
the number of joint epilogues that must have completed before its MMAs may be issued.  These tests
re-derive the hazards from the table itself - shared-memory tensors (who wrote what a job reads),
accumulator slots (who drained the slot a job overwrites), the one-barrier-per-epilogue arrays - so that an edit
of the job order or of the buffer layout that forgets a dependency fails here, on the CPU."""
import numpy as np
import pytest

from conftest import MODELS, model_path

FIELDS = ('n idesc ntiles L lp ntaps tap0 tap1 tap2 lo16 ncb cb0 w_goff tcol wp0 wp1 first last kind bias bn '
          'out_off out_lp out_lo out_cg out_ncg out_L edge15 zero_y joint need eseq').split()
EPI_N48, EPI_N48_POOL_BN, EPI_N48_BN, EPI_N16, EPI_PARITY, EPI_HEAD = range(6)
JOINT_NONE, JOINT_PAIR, JOINT_STACK = range(3)
W_PART0, W_PART1 = 15360, 12288
RING = 16     # kJointRing: one MMA-done / epilogue-done / weights barrier per joint epilogue resp. joint job, never reused


def job_table(name, which):
    from deepbinner_b200 import _native, weights
    lib = _native.load_library()
    blob = weights.load_blob(model_path(name))
    out = np.zeros((32, 32), np.int32)
    n = lib.db_tc_job_table(blob, len(blob), which, _native.as_ptr(out), 32)
    assert n > 0, lib.db_last_error()
    return [dict(zip(FIELDS, (int(v) for v in row))) for row in out[:n]]


def in_extent(job):
    """Byte interval (relative to the window's region) of the hi array the job's MMAs read."""
    first_row = job['tap0'] * 16 - (0 if job['ntaps'] == 3 else 16)
    ngroups = 2 * job['ncb']
    return first_row, first_row + ngroups * job['lp'] * 16


def out_extent(job):
    return job['out_off'], job['out_off'] + 2 * job['out_lo']


@pytest.mark.parametrize('which,njobs', [(0, 21)])
@pytest.mark.parametrize('name', MODELS)
def test_joint_schedule_is_hazard_free(name, which, njobs):
    jobs = job_table(name, which)
    solo = which == 1
    assert len(jobs) == njobs
    joint = [j for j in jobs if j['joint'] != JOINT_NONE]
    assert jobs.index(joint[0]) + len(joint) == len(jobs), 'joint jobs form the tail of the table'
    assert all(j['first'] and j['last'] for j in jobs if j['joint'] == JOINT_NONE)

    for j in jobs:
        nkb = j['ntaps'] * j['ncb']
        assert nkb in (3, 9) and j['w_goff'] % 128 == 0
        if solo:   # K-block major: wp0 = bytes of one K block ((hi, lo) x 2 chunks x n rows x 16 B), wp1 = K blocks
            assert j['wp0'] == 2 * 2 * j['n'] * 16 <= 3072 and j['wp1'] == nkb
        else:      # weights of every job fit the two-part buffer; every part is a whole number of 16-byte rows
            assert 0 < j['wp0'] <= W_PART0 and 0 < j['wp1'] <= W_PART1
            assert j['wp0'] % 16 == 0 and j['wp1'] % 16 == 0
            assert j['wp0'] + j['wp1'] == nkb * 2 * 2 * j['n'] * 16   # K blocks x (hi, lo) x 2 chunks x n rows

    # hand-off sequence: the joint jobs (pair kernel) / every job, after conv1d_1's stage = epilogue 0 (solo)
    seq = jobs if solo else joint
    e0 = 1 if solo else 0
    # epilogue sequence numbers: consecutive over the jobs that have an epilogue; need never decreases
    eseq = [j['eseq'] for j in seq if j['last']]
    assert eseq == list(range(e0, e0 + len(eseq)))
    assert all(j['eseq'] == -1 for j in seq if not j['last'])
    needs = [j['need'] for j in seq]
    assert needs == sorted(needs) and needs[0] == e0
    if solo:   # conv1d_2 .. conv1d_9 form a chain: each waits for every earlier epilogue, columns from 0
        assert all(j['need'] == j['eseq'] and j['tcol'] == 0 for j in jobs if j['joint'] == JOINT_NONE)
        assert all(j['tcol'] + 64 <= 256 for j in jobs) and all(j['ntiles'] <= 4 for j in jobs)

    slot_drained_by = {}     # accumulator slot -> eseq of the epilogue that last read it
    writer_of = []           # (extent, eseq, channel groups of the concat tensor or None) written by epilogues of the sequence
    done_epilogues = e0
    for k, j in enumerate(seq):
        # (a) the tensor the job reads was written by an epilogue that `need` covers
        if j['joint'] == JOINT_STACK and j['kind'] == EPI_N48_BN and j['ntaps'] == 3 and j['ncb'] == 3 and \
                j['lo16'] * 16 > 8192:
            # a K-slice of conv1d_17 reads channel groups [2 cb0, 2 cb0 + 2 ncb) of the concat tensor: the branch
            # whose epilogue wrote exactly those groups must be complete
            g0, g1 = 2 * j['cb0'], 2 * j['cb0'] + 2 * j['ncb']
            producers = [e for (_, e, parity) in writer_of if parity is not None and parity[0] < g1 and g0 < parity[1]]
            assert producers, (k, j)
        else:
            lo, hi = in_extent(j)
            producers = [e for ((a, b), e, parity) in writer_of if parity is None and a < hi and lo < b]
        if producers:
            assert j['need'] >= max(producers) + 1, (k, j, producers)
        # (b) the accumulator slot(s) were drained
        cols = [j['tcol'] + 64 * t for t in range(j['ntiles'])]
        for c in cols:
            if j['first'] and c in slot_drained_by:
                assert j['need'] >= slot_drained_by[c] + 1, (k, j)
        # (c) every joint job / joint epilogue has its own barrier
        if j['last']:
            done_epilogues += 1
        assert done_epilogues <= RING and k < RING, (k, j)
        if j['last']:
            for c in cols:
                slot_drained_by[c] = j['eseq']
            if j['kind'] != EPI_HEAD:
                groups = (j['out_cg'], j['out_cg'] + j['out_ncg']) if j['kind'] == EPI_PARITY else None
                writer_of.append((out_extent(j), j['eseq'], groups))
    # the head is the last epilogue and waits for everything before it
    assert joint[-1]['kind'] == EPI_HEAD and joint[-1]['need'] == eseq[-1]


@pytest.mark.parametrize('name', MODELS)
def test_parameter_blocks_fit(name):
    for which, cap in ((0, 1664),):
        jobs = job_table(name, which)
        used = max(max(j['bias'] + j['n'], j['bn'] + 96 if j['bn'] else 0) for j in jobs)
        assert used <= cap


# ---------------------------------------------------------------------------------------------
# packed operands: what the jobs feed the tensor cores with, against the model's own tensors
# ---------------------------------------------------------------------------------------------
# job -> (conv layers computed by it (second = appended along N), BatchNorm applied in its epilogue)
FUSED_LAYERS = [((2,), 0), ((3,), 0), ((4,), 2), ((5,), 0), ((6,), 0), ((7,), 3), ((8,), 0), ((9,), 4),
                ((12, 14), 0), ((10,), 5), ((15,), 0), ((11,), 5), ((13,), 5), ((16,), 5),
                ((17,), 0), ((17,), 0), ((17,), 0), ((17,), 6), ((18,), 0), ((19,), 7), ((20,), 0)]


def packed(name, which):
    import ctypes
    from deepbinner_b200 import _native, weights
    lib = _native.load_library()
    blob = weights.load_blob(model_path(name))
    wb, pf = ctypes.c_int64(), ctypes.c_int64()
    assert lib.db_tc_packed(blob, len(blob), which, None, 0, None, 0, ctypes.byref(wb), ctypes.byref(pf)) == 0
    w = np.zeros(wb.value, np.uint8)
    prm = np.zeros(pf.value, np.float32)
    assert lib.db_tc_packed(blob, len(blob), which, _native.as_ptr(w), len(w), _native.as_ptr(prm), len(prm),
                            ctypes.byref(wb), ctypes.byref(pf)) == 0
    return w, prm


def bf16_pairs_to_f32(buf):
    return (buf.view(np.uint16).astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize('which,layers', [(0, FUSED_LAYERS)])
@pytest.mark.parametrize('name', MODELS)
def test_packed_weights_and_parameters_match_the_model(name, which, layers):
    from oracle import deepbinner_oracle as orc
    model = orc.load_weights(model_path(name), dtype=np.float64)
    jobs = job_table(name, which)
    w, prm = packed(name, which)
    assert len(jobs) == len(layers)
    for j, (convs, bn) in zip(jobs, layers):
        n, ntaps, ncb, cb0 = j['n'], j['ntaps'], j['ncb'], j['cb0']
        nkb = ntaps * ncb
        split = 5 if nkb == 9 else 2
        # expected kernel [ntaps][cin][n]: layers appended along N, average pool folded (W/3 on three taps)
        kernels = [model['conv1d_%d/kernel' % c] for c in convs]
        folded = j['edge15'] == 1
        assert folded == (convs == (10,))
        cin_total = kernels[0].shape[1]
        expect = np.zeros((ntaps, cin_total, n))
        col = 0
        for k in kernels:
            kk = np.repeat(k / 3.0, 3, axis=0) if folded else k
            assert kk.shape[0] == ntaps and kk.shape[1] == cin_total
            expect[:, :, col:col + kk.shape[2]] = kk
            col += kk.shape[2]
        # unpack [part 0 | part 1], part = [hi blocks | lo blocks], block = [2 chunks][n rows][8]
        # (solo kernel: one part per K block)
        got = np.zeros((ntaps, cin_total, n))
        seen = np.zeros((ntaps, cin_total), bool)
        off = j['w_goff']
        parts = [(kb, kb + 1) for kb in range(nkb)] if which == 1 else [(0, split), (split, nkb)]
        for kb0, kb1 in parts:
            cnt = kb1 - kb0
            part = bf16_pairs_to_f32(w[off:off + 2 * cnt * 2 * n * 8 * 2]).astype(np.float64)
            hi, lo = part[:cnt * 2 * n * 8], part[cnt * 2 * n * 8:]
            both = (hi + lo).reshape(cnt, 2, n, 8)
            assert np.all(np.abs(lo) <= np.abs(hi) * 2.0 ** -7 + 1e-30)      # lo is the remainder of hi
            for kb in range(kb0, kb1):
                t, cb = divmod(kb, ncb)
                for chunk in range(2):
                    c0 = (cb0 + cb) * 16 + chunk * 8
                    got[t, c0:c0 + 8, :] = both[kb - kb0, chunk].T
                    seen[t, c0:c0 + 8] = True
            off += 2 * cnt * 2 * n * 8 * 2
        assert off - j['w_goff'] == (j['wp0'] * j['wp1'] if which == 1 else j['wp0'] + j['wp1'])
        scale = np.abs(expect).max()
        assert np.abs(got[seen] - expect[seen]).max() <= scale * 2.0 ** -15, (j, convs)
        if convs != (17,):
            assert seen.all()
        else:
            assert seen[:, cb0 * 16:(cb0 + 3) * 16].all() and seen.sum() == 3 * 48   # one K-slice of conv1d_17
        if not j['last']:
            continue
        # bias (zero padded to n) and folded BatchNorm of the epilogue
        bias = np.concatenate([model['conv1d_%d/bias' % c] for c in convs])
        got_bias = prm[j['bias']:j['bias'] + n]
        assert np.allclose(got_bias[:len(bias)], bias, rtol=1e-6, atol=0) and not got_bias[len(bias):].any()
        assert (bn != 0) == (j['bn'] != 0)
        if bn:
            p = 'batch_normalization_%d/' % bn
            s = model[p + 'gamma'] / np.sqrt(model[p + 'moving_variance'] + 1e-3)
            h = model[p + 'beta'] - model[p + 'moving_mean'] * s
            ch0 = j['out_cg'] * 8 if j['kind'] == EPI_PARITY else 0
            assert np.allclose(prm[j['bn']:j['bn'] + 48], s[ch0:ch0 + 48], rtol=1e-6)
            assert np.allclose(prm[j['bn'] + 48:j['bn'] + 96], h[ch0:ch0 + 48], rtol=1e-5, atol=1e-6)
