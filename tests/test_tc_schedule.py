"""Host-side checks of the tcgen05 engine's MMA job tables (csrc/dbn_tc.cu build_jobs /
build_tail_jobs), dumped through the C ABI without a GPU (db_tc_job_table).

The joint phase lets the MMA issuer run ahead of the epilogue warps: every joint job carries `need`,
the number of joint epilogues that must have completed before its MMAs may be issued.  These tests
re-derive the hazards from the table itself - shared-memory tensors (who wrote what a job reads),
accumulator slots (who drained the slot a job overwrites), the 4-deep mbarrier rings - so that an edit
of the job order or of the buffer layout that forgets a dependency fails here, on the CPU."""
import numpy as np
import pytest

from conftest import MODELS, model_path

FIELDS = ('n idesc ntiles L lp ntaps tap0 tap1 tap2 lo16 ncb cb0 w_goff tcol wp0 wp1 first last kind bias bn '
          'out_off out_lp out_lo out_cg out_ncg out_L edge15 zero_y joint need eseq').split()
EPI_N48, EPI_N48_POOL_BN, EPI_N48_BN, EPI_N16, EPI_PARITY, EPI_HEAD = range(6)
JOINT_NONE, JOINT_PAIR, JOINT_STACK = range(3)
W_PART0, W_PART1 = 15360, 12288
RING = 4


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


@pytest.mark.parametrize('which,njobs', [(0, 21), (1, 18)])
@pytest.mark.parametrize('name', MODELS)
def test_joint_schedule_is_hazard_free(name, which, njobs):
    jobs = job_table(name, which)
    assert len(jobs) == njobs
    joint = [j for j in jobs if j['joint'] != JOINT_NONE]
    assert jobs.index(joint[0]) + len(joint) == len(jobs), 'joint jobs form the tail of the table'
    assert all(j['first'] and j['last'] for j in jobs if j['joint'] == JOINT_NONE)

    # weights of every job fit the two-part buffer; every part is a whole number of 16-byte rows
    for j in jobs:
        assert 0 < j['wp0'] <= W_PART0 and 0 < j['wp1'] <= W_PART1
        assert j['w_goff'] % 128 == 0 and j['wp0'] % 16 == 0 and j['wp1'] % 16 == 0
        nkb = j['ntaps'] * j['ncb']
        assert nkb in (3, 9)
        assert j['wp0'] + j['wp1'] == nkb * 2 * 2 * j['n'] * 16   # K blocks x (hi, lo) x 2 chunks x n rows

    # epilogue sequence numbers: consecutive over the jobs that have an epilogue; need never decreases
    eseq = [j['eseq'] for j in joint if j['last']]
    assert eseq == list(range(len(eseq)))
    assert all(j['eseq'] == -1 for j in joint if not j['last'])
    needs = [j['need'] for j in joint]
    assert needs == sorted(needs) and needs[0] == 0

    slot_drained_by = {}     # accumulator slot -> eseq of the epilogue that last read it
    writer_of = []           # (extent, eseq, is_parity) of tensors written by joint epilogues
    done_epilogues = 0
    for k, j in enumerate(joint):
        # (a) the tensor the job reads was written by an epilogue that `need` covers
        if j['joint'] == JOINT_STACK and j['kind'] == EPI_N48_BN and j['ntaps'] == 3 and j['ncb'] == 3 and \
                j['lo16'] * 16 > 8192:
            producers = [e for (_, e, parity) in writer_of if parity]          # conv1d_17 reads the concat tensor
        else:
            lo, hi = in_extent(j)
            producers = [e for ((a, b), e, parity) in writer_of if not parity and a < hi and lo < b]
        if producers:
            assert j['need'] >= max(producers) + 1, (k, j, producers)
        # (b) the accumulator slot was drained
        if j['first'] and j['tcol'] in slot_drained_by:
            assert j['need'] >= slot_drained_by[j['tcol']] + 1, (k, j)
        # (c) the 4-deep mbarrier rings never hold more than three unconsumed phases
        if j['last']:
            done_epilogues += 1
        assert done_epilogues - j['need'] <= RING - 1, (k, j)
        if j['last']:
            slot_drained_by[j['tcol']] = j['eseq']
            if j['kind'] != EPI_HEAD:
                writer_of.append((out_extent(j), j['eseq'], j['kind'] == EPI_PARITY))
    # the head is the last epilogue and waits for everything before it
    assert joint[-1]['kind'] == EPI_HEAD and joint[-1]['need'] == len(eseq) - 1


@pytest.mark.parametrize('name', MODELS)
def test_parameter_blocks_fit(name):
    for which, cap in ((0, 1664), (1, 1424)):
        jobs = job_table(name, which)
        used = max(max(j['bias'] + j['n'], j['bn'] + 96 if j['bn'] else 0) for j in jobs)
        assert used <= cap
