"""Discrete-event model of the mbarrier protocol of the tensor-core kernel (csrc/dbn_tc.cu, k_tc_forward), run on the
CPU against the REAL job table (db_tc_job_table) under randomised interleavings.

The kernel's four roles - the epilogue warps, the two MMA issuers, the weight loader - loop over the window pairs of a
persistent CTA on their own and meet only through mbarriers that are never re-initialised: every wait passes a phase
PARITY derived from the job index / the pair index, tcgen05.commit arrives when the committing thread's earlier MMAs have
completed, cp.async.bulk completes a transaction count.  This model restates each role's sequence of waits, arrivals,
commits and copies exactly as the kernel performs them and checks, over several pairs per CTA and many schedules:

  * no role ever blocks for good (deadlock), every role finishes every pair;
  * a parity wait returns exactly when the phase the role MEANS has completed - never a phase early (parity aliasing
    after a skipped phase) and never one late;
  * every completed phase of every barrier was waited for by somebody (what compute-sanitizer's synccheck demands);
  * data hazards: MMAs are issued only after their input was written and their accumulator cells were drained, an epilogue
    only runs after all its MMAs completed, a weight slot is only overwritten when no issued MMA still reads it and only
    read when the right job's weights have landed in it.

It is a model, not the product: the GPU suite (tests/test_gpu_tc_layers.py, the sanitizer runs under profiles/) checks the
kernel itself; this file makes an edit of the job order, the owner assignment or a parity rule fail on the CPU first."""
import random

import pytest

from conftest import MODELS
from test_tc_schedule import JOINT_NONE, job_table

N_SINGLE = 8   # conv1d_2 .. conv1d_9: one pass per window


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, count, 0
        self.waited = set()          # phases somebody waited for

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, self.name
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count

    def ready(self, parity):         # mbarrier.try_wait.parity: the phase with this parity has completed
        return (self.completed & 1) != parity


class Cta:
    """State shared by the roles of one CTA."""

    def __init__(self, jobs, npairs, rng):
        self.jobs, self.npairs, self.rng = jobs, npairs, rng
        self.joint = [j for j in jobs if j['joint'] != JOINT_NONE]
        njoint, nepi = len(self.joint), sum(1 for j in self.joint if j['last'])
        b = {}
        for p in (0, 1):
            b['wfull', p] = Barrier('wfull%d' % p, 2)     # expect_tx arrival + the bulk copy's completion
            b['wfree', p] = Barrier('wfree%d' % p, 2)     # one commit per window pass
            b['mma', p] = Barrier('mma%d' % p, 1)
            b['epi', p] = Barrier('epi%d' % p, 1)         # (384 arrivals in the kernel: the epilogue threads move as one here)
        b['final'] = Barrier('final', 2)
        b['x'] = Barrier('x', 2)
        for k in range(njoint):
            b['jwfull', k] = Barrier('jwfull%d' % k, 2)
            b['jwfree', k] = Barrier('jwfree%d' % k, 1)
        for e in range(nepi):
            b['jmma', e] = Barrier('jmma%d' % e, 1)
            b['jepi', e] = Barrier('jepi%d' % e, 1)
        self.bar = b
        # issuer assignment as tc_create() does it
        owner, jk = 1, 0
        for j in jobs:
            j['both'] = j['joint'] == JOINT_NONE and j['L'] >= 512
            if j['joint'] != JOINT_NONE:
                if j['first']:
                    owner ^= 1
                j['owner'], j['jk'] = owner, jk
                jk += 1
        self.fifo = [[], []]              # per issuer: MMAs / commits in issue order ("the tensor pipe")
        self.copies = []                  # bulk copies in flight
        self.mma_done = set()
        self.epi_done = set()
        self.slot_content = {}            # weight slot -> (pair, job) whose weights have landed
        self.slot_readers = {}            # weight slot -> MMA groups issued and not yet completed
        self.tmem_busy = {}               # accumulator cell (64 columns) -> job whose epilogue has not drained it yet

    # ---- asynchronous completions ----
    def pipe_step(self, i):
        kind, what = self.fifo[i].pop(0)
        if kind == 'mma':
            key, slots = what
            self.mma_done.add(key)
            for s in slots:
                self.slot_readers[s].discard(key)
        else:
            what.arrive()

    def copy_step(self, k):
        slot, content, bar = self.copies.pop(k)
        self.slot_content[slot] = content
        bar.arrive()

    # ---- helpers used by the roles ----
    def start_copy(self, slot, content, bar):
        assert not self.slot_readers.get(slot), ('weight slot overwritten while MMAs read it', slot, content)
        self.slot_content[slot] = None
        bar.arrive()                      # mbarrier.arrive.expect_tx
        self.copies.append((slot, content, bar))

    def issue(self, i, key, slots, content):
        for s in slots:
            assert self.slot_content.get(s) == content, ('MMA reads a weight slot that does not hold its job', s, key)
            self.slot_readers.setdefault(s, set()).add(key)
        self.fifo[i].append(('mma', (key, slots)))

    def commit(self, i, bar):
        self.fifo[i].append(('commit', bar))


def wait(bar, parity, expect):
    """Yielded by a role: block until the parity wait succeeds; `expect` = the absolute phase the role means."""
    return ('wait', bar, parity, expect)


def shadow(bar, parity, expect):
    """A wait whose only purpose is to keep a role in step with a barrier it will need later (issuer 1 during
    conv1d_2..4).  Nothing holds the barrier back for such a waiter, so the kernel relies on it being PROMPT: a dozen
    instructions per job against >= 4 000 cycles per phase.  The model serves these waits before anything else happens
    (`prompt_shadows`); test_shadow_waits_rely_on_promptness shows what the checks report without that."""
    return ('shadow', bar, parity, expect)


def cells_of(j, w):
    if j['joint'] != JOINT_NONE:
        return [j['tcol'] // 64]
    return [4 * w + t for t in range(j['ntiles'])]


def joint_slots(jk):
    s = (jk + 1) % 3
    return ['w0', 'w1'] if s == 0 else ['js%d' % s]


def epilogue_role(c):
    b, jobs = c.bar, c.jobs
    for it in range(c.npairs):
        ph = it & 1
        for w in (0, 1):                                   # z-score + conv1d_1
            yield None
            c.epi_done.add(('conv1', it, w))
            b['epi', w].arrive()
        mma_phase = [0, 0]
        for j, J in enumerate(jobs):
            if not J['last']:
                continue
            if J['joint'] != JOINT_NONE:
                e = J['eseq']
                yield wait(b['jmma', e], ph, it)
                chain = j
                while not jobs[chain]['first']:
                    chain -= 1
                for k in range(chain, j + 1):
                    assert ('jm', it, k) in c.mma_done, ('joint epilogue before its MMAs completed', it, j, k)
                yield None
                for cell in cells_of(J, 0):
                    assert c.tmem_busy.pop(cell) == (it, chain)
                c.epi_done.add(('jepi', it, e))
                if J['last'] & 2:
                    b['jepi', e].arrive()
                continue
            for w in (0, 1):
                yield wait(b['mma', w], mma_phase[w], it * N_SINGLE + j)
                mma_phase[w] ^= 1
                assert ('m', it, j, w, 0) in c.mma_done and ('m', it, j, w, 1) in c.mma_done
                yield None
                for cell in cells_of(J, w):
                    assert c.tmem_busy.pop(cell) == (it, j, w)
                c.epi_done.add(('epi', it, j, w))
                if j + 1 < len(jobs) and jobs[j + 1]['joint'] != JOINT_NONE:
                    b['x'].arrive()
                else:
                    b['epi', w].arrive()


def issuer_role(c, me):
    b, jobs = c.bar, c.jobs
    for it in range(c.npairs):
        ph = it & 1
        in_joint = False
        for j, J in enumerate(jobs):
            par = j & 1
            if not (J['both'] and me != 0):
                yield None
            if J['joint'] != JOINT_NONE:
                if J['owner'] != me:
                    continue
                jk = J['jk']
                yield wait(b['jwfull', jk], ph, it)
                if not in_joint:
                    yield wait(b['x'], ph, it)
                in_joint = True
                if J['need'] > 0:
                    yield wait(b['jepi', J['need'] - 1], ph, it)
                # inputs: the last single-window epilogues of both windows, every joint epilogue below `need`
                for w in (0, 1):
                    assert ('epi', it, N_SINGLE - 1, w) in c.epi_done
                for e in range(J['need']):
                    assert ('jepi', it, e) in c.epi_done, ('joint job issued before a needed epilogue', it, j, e)
                if J['first']:
                    for cell in cells_of(J, 0):
                        assert cell not in c.tmem_busy, ('accumulator cell overwritten before it was drained', it, j, cell)
                        c.tmem_busy[cell] = (it, j)
                c.issue(me, ('jm', it, j), joint_slots(jk), (it, j))
                if J['last']:
                    c.commit(me, b['jmma', J['eseq']])
                c.commit(me, b['jwfree', jk])
                continue
            if J['both'] and me != 0:      # issuer 1 only follows the barriers it will wait on later
                yield shadow(b['epi', 1], par, it * N_SINGLE + j)
                yield shadow(b['wfull', 0], par, it * N_SINGLE + j)
                yield shadow(b['wfull', 1], par, it * N_SINGLE + j)
                continue
            windows = (0, 1) if J['both'] else (me,)
            nfree = 2 if J['both'] else 1
            for w in windows:
                free_now = w == windows[-1]
                if w == windows[0]:
                    yield wait(b['wfull', 0], par, it * N_SINGLE + j)
                yield wait(b['epi', w], par, it * N_SINGLE + j)
                prev = ('conv1', it, w) if j == 0 else ('epi', it, j - 1, w)
                assert prev in c.epi_done, ('MMAs issued before their input was written', it, j, w)
                for cell in cells_of(J, w):
                    assert cell not in c.tmem_busy, ('accumulator cell overwritten before it was drained', it, j, w, cell)
                    c.tmem_busy[cell] = (it, j, w)
                c.issue(me, ('m', it, j, w, 0), ['w0'], (it, j))
                if free_now:
                    for _ in range(nfree):
                        c.commit(me, b['wfree', 0])
                yield None
                if w == windows[0]:
                    yield wait(b['wfull', 1], par, it * N_SINGLE + j)
                c.issue(me, ('m', it, j, w, 1), ['w1'], (it, j))
                if J['last']:
                    c.commit(me, b['mma', w])
                if free_now:
                    for _ in range(nfree):
                        c.commit(me, b['wfree', 1])
    c.commit(me, b['final'])
    yield wait(b['final'], 0, 0)


def loader_role(c):
    b, jobs = c.bar, c.jobs
    for it in range(c.npairs):
        ph = it & 1
        free_phase, jk, nfree_waits = 0, 0, 0
        for j, J in enumerate(jobs):
            yield None
            if J['joint'] != JOINT_NONE:
                if jk == 2:
                    if j > 2:
                        yield wait(b['wfree', 0], free_phase, it * N_SINGLE + nfree_waits)
                        yield wait(b['wfree', 1], free_phase, it * N_SINGLE + nfree_waits)
                elif jk >= 3:
                    yield wait(b['jwfree', jk - 3], ph, it)
                slots = joint_slots(jk)
                # one bulk copy fills the slot; a two-part slot (the normal weight buffer) is modelled as its two halves
                for s in slots[:-1]:
                    assert not c.slot_readers.get(s)
                    c.slot_content[s] = (it, j)
                c.start_copy(slots[-1], (it, j), b['jwfull', jk])
                jk += 1
                continue
            if j > 0:
                yield wait(b['wfree', 0], free_phase, it * N_SINGLE + nfree_waits)
            c.start_copy('w0', (it, j), b['wfull', 0])
            if j > 0:
                yield wait(b['wfree', 1], free_phase, it * N_SINGLE + nfree_waits)
                free_phase ^= 1
                nfree_waits += 1
            c.start_copy('w1', (it, j), b['wfull', 1])
        for k in range(max(jk - 3, 0), jk):
            yield wait(b['jwfree', k], ph, it)


def run(jobs, npairs, seed, prompt_shadows=True):
    rng = random.Random(seed)
    c = Cta([dict(j) for j in jobs], npairs, rng)
    roles = {'epilogue': epilogue_role(c), 'issuer0': issuer_role(c, 0), 'issuer1': issuer_role(c, 1),
             'loader': loader_role(c)}
    blocked = {}     # role -> [request, phases completed when the wait succeeded (None: still waiting)]
    # a role-specific pace makes whole classes of schedules likely (a slow loader, a slow epilogue, ...)
    pace = {name: rng.choice((1, 1, 3, 10)) for name in list(roles) + ['pipe0', 'pipe1', 'copy']}

    def resume(name):
        if name in blocked:
            (_, bar, parity, expect), seen = blocked.pop(name)
            assert seen == expect + 1, ('%s: wait on %s parity %d succeeded with %d phases completed, meant phase %d'
                                        % (name, bar.name, parity, seen, expect))
            bar.waited.add(expect)
        try:
            req = next(roles[name])
        except StopIteration:
            del roles[name]
            return
        if req is not None:
            blocked[name] = [req, None]

    def latch():
        # a thread blocked in mbarrier.try_wait is woken when the phase completes: the wait has succeeded at that
        # moment, however late the thread runs again
        for entry in blocked.values():
            (_, bar, parity, _), seen = entry
            if seen is None and bar.ready(parity):
                entry[1] = bar.completed

    steps = 0
    while roles:
        steps += 1
        assert steps < 2_000_000
        latch()
        if prompt_shadows:
            prompt = [n for n, (req, seen) in blocked.items() if req[0] == 'shadow' and seen is not None]
            if prompt:
                resume(prompt[0])
                continue
        choices = []
        for name in roles:
            if name in blocked and blocked[name][1] is None:
                continue
            choices += [('role', name)] * pace[name]
        for i in (0, 1):
            if c.fifo[i]:
                choices += [('pipe', i)] * pace['pipe%d' % i]
        for k in range(len(c.copies)):
            choices += [('copy', k)] * pace['copy']
        assert choices, 'deadlock: ' + ', '.join('%s waits for %s parity %d (completed %d, means phase %d)' % (
            n, w[0][1].name, w[0][2], w[0][1].completed, w[0][3]) for n, w in blocked.items())
        kind, what = rng.choice(choices)
        if kind == 'pipe':
            c.pipe_step(what)
        elif kind == 'copy':
            c.copy_step(what)
        else:
            resume(what)
    assert not c.fifo[0] and not c.fifo[1] and not c.copies and not c.tmem_busy
    for bar in c.bar.values():
        missing = set(range(bar.completed)) - bar.waited
        assert not missing, ('%s: phases %s completed without a waiter' % (bar.name, sorted(missing)))
        assert bar.pending == bar.count, (bar.name, 'ends in the middle of a phase')
    return steps


@pytest.mark.parametrize('npairs', [1, 2, 3, 4])
def test_barrier_protocol_under_random_schedules(npairs):
    jobs = job_table(MODELS[0], 0)
    assert sum(1 for j in jobs if j['joint'] == JOINT_NONE) == N_SINGLE
    for seed in range(40):
        run(jobs, npairs, 1000 * npairs + seed)


def test_shadow_waits_rely_on_promptness():
    """The one timing assumption of the protocol, made explicit: while issuer 0 issues conv1d_2..4 for both windows, issuer 1
    only follows bar_epi[1] and the weight barriers so that its parities are right when conv1d_5 starts.  Nothing makes
    those barriers wait for issuer 1; if it could fall two phases behind, its parity waits would alias - which is what
    the model reports as soon as the shadow waits are scheduled like everything else."""
    jobs = job_table(MODELS[0], 0)
    failures = 0
    for seed in range(40):
        try:
            run(jobs, 2, seed, prompt_shadows=False)
        except AssertionError as e:
            assert 'issuer1' in str(e) or 'deadlock' in str(e), e
            failures += 1
    assert failures > 0


def test_the_model_notices_a_broken_protocol():
    """The checks above are not vacuous: three deliberate protocol errors are each caught."""
    jobs = job_table(MODELS[0], 0)
    joint0 = next(i for i, j in enumerate(jobs) if j['joint'] != JOINT_NONE)

    def broken(mutate, pairs=2):
        bad = [dict(j) for j in jobs]
        mutate(bad)
        for seed in range(40):
            try:
                run(bad, pairs, seed)
            except AssertionError:
                return True
        return False

    # a joint job that does not wait for the epilogue that writes its input
    def need_too_small(bad):
        k = max(i for i, j in enumerate(bad) if j['need'] > 1)
        bad[k]['need'] = 0
        for j in bad[joint0:]:      # (the kernel only waits for epilogue need - 1: keep the arrival flags consistent)
            j['last'] |= 2 if j['last'] else 0
    assert broken(need_too_small)

    # an epilogue that nobody waits for still arrives on its barrier (the synccheck finding of round 2)
    def arrival_without_waiter(bad):
        for j in bad[joint0:]:
            if j['last'] == 1:
                j['last'] = 3
                return
    assert broken(arrival_without_waiter)

    # two consecutive joint jobs share an accumulator slot although the first has not been drained
    def slot_reused_too_early(bad):
        firsts = [i for i, j in enumerate(bad) if j['joint'] != JOINT_NONE and j['first']]
        bad[firsts[1]]['tcol'] = bad[firsts[0]]['tcol']
    assert broken(slot_reused_too_early)
