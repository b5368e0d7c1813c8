#!/usr/bin/env python3
"""
Benchmark of the classification hot path (BASELINE.json metric: reads classified/sec on synthetic
1024-sample float32 signal windows at batch 256; softmax max-abs-err vs the CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--engine fp32|tcgen05]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config "synthetic 1M reads x 1024 samples, EXP-NBD103 start model"; SURVEY 8d
config 4): every rank owns a resident shard of SHARD synthetic 1024-sample reads (scan_size 512 =>
one window per read): the reference's own gaussian random-signal recipe (balance.py:171-174) with
10 % of the windows cut from the reference's real fast5 fixtures, z-scored as call_batch does.
One "step" = one pass of the network over the whole shard in launches of batch 256.
The shard (256 MiB of fp32) is larger than the 126 MB L2, so no flush is needed between steps.

value  = reads/s, whole job, inputs already in HBM (device-resident entry of the C ABI),
         CUDA-event timed, max over ranks.
e2e    = the same metric through the reference-facing entry - the fused `call_batch` (C-ABI
         db_call_batch_submit_packed / _wait; what classify.call_batch runs) on the raw int16 reads in
         HOST memory, scan_size 512: gather into pinned staging, H2D of the samples and D2H of the
         per-read probabilities and calls inside the timed region.  `e2e.predict_f32` is the same through
         `B200Model.predict` on float32 windows (seam b1), `e2e.predict_n256_f64` the shape the reference's
         own call_batch produces (pageable float64, 256 windows per call).
roofline = tensor (dense conv contraction): algorithmic 33,629,952 FLOP per window, against the measured
         burst bf16 peak (`frac`) and the sustained one (`frac_sustained`).
cpu_baseline = the torch-CPU fp32 oracle on all host cores, bounded sample (the reference's own
         TensorFlow-CPU model.predict cannot run in this image - no tensorflow/keras/h5py).
configs = first-class lines for BASELINE configs[1] (EXP-NBD103 start+end, batch 256), configs[2]
         (SQK-RBK004, batch 512) and configs[4] (realtime streaming, start+end, rounds of 20 000 reads per
         GPU) through the product's own batch pipeline (classify.classify_read_batches).

`--config strong` runs the strong-scaling form of configs[3]: 1 048 576 reads in total, split over the ranks.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_PER_WINDOW = 33629952          # 2 x 16,814,976 MACs
BATCH = 256                          # launch batch of the metric
SHARD = 65536                        # reads resident per GPU and processed per step
MODEL = 'EXP-NBD103_read_starts'
N_STREAMS = int(os.environ.get('DBN_BENCH_STREAMS', '4'))   # launches of the headline loop rotate over this many streams


def model_path():
    return str(ROOT / 'deepbinner_b200' / 'models' / (MODEL + '.dbnw'))


def measured_peaks():
    """(burst, sustained, source): the timed region is short (well under a second at full clocks, no
    power-cap samples), so the burst figure is the roofline denominator; sustained is reported beside it."""
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        burst = d.get('bf16_tflops')
        return burst, d.get('bf16_tflops_sustained', burst), 'MEASURED_PEAKS.json bf16_tflops (burst; measured)'
    return 1590.0, 1400.0, 'fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained)'


def synthetic_reads(n, seed, length=1024):
    """int16 reads of the reference's gaussian recipe (balance.py:171-174) with 10 % real-fixture cuts:
    the raw-signal form of synthetic_windows (which z-scores the same values)."""
    rng = np.random.RandomState(seed)
    z = np.load(ROOT / 'tests' / 'golden' / 'fixture_reads.npz')
    real = [z['signal_{}'.format(i)] for i in range(7)]
    out = np.empty((n, length), dtype=np.int16)
    block = 8192
    for s in range(0, n, block):
        m = min(block, n - s)
        mean = rng.uniform(300, 600, (m, 1))
        sd = rng.uniform(10, 500, (m, 1))
        x = np.trunc(rng.standard_normal((m, length)) * sd + mean)
        for j in range(m // 10):
            sig = real[rng.randint(7)]
            if len(sig) > length:
                a = rng.randint(0, len(sig) - length)
                x[j * 10] = sig[a:a + length]
        out[s:s + m] = np.clip(x, -32768, 32767).astype(np.int16)
    return out


class PackedReads:
    """Reads of equal length in one int16 buffer, in the form load_fast5s.PackedSignals has (what the
    native fast5 reader hands to call_batch): `samples`, `offsets`, `rows`."""

    def __init__(self, reads2d):
        n, length = reads2d.shape
        self.samples = np.ascontiguousarray(reads2d).reshape(-1)
        self.offsets = np.arange(n + 1, dtype=np.int64) * length
        self.rows = np.arange(n, dtype=np.int64)
        self.n = n

    def __len__(self):
        return self.n


def synthetic_windows(n, seed):
    """float32 [n, 1024] z-scored windows: gaussian recipe of reference balance.py:171-174 (per read
    mean~U(300,600), sd~U(10,500), int(N(mean,sd))) + 10 % windows cut from the real fixtures."""
    rng = np.random.RandomState(seed)
    z = np.load(ROOT / 'tests' / 'golden' / 'fixture_reads.npz')
    real = [z['signal_{}'.format(i)] for i in range(7)]
    out = np.empty((n, 1024), dtype=np.float32)
    block = 8192
    for s in range(0, n, block):
        m = min(block, n - s)
        mean = rng.uniform(300, 600, (m, 1))
        sd = rng.uniform(10, 500, (m, 1))
        x = np.trunc(rng.standard_normal((m, 1024)) * sd + mean)
        n_real = m // 10
        for j in range(n_real):
            sig = real[rng.randint(7)]
            a = rng.randint(0, len(sig) - 1024)
            x[j * 10] = sig[a:a + 1024]
        mu = x.mean(axis=1, keepdims=True)
        sg = x.std(axis=1, keepdims=True)
        sg[sg == 0] = 1.0
        out[s:s + m] = ((x - mu) / sg).astype(np.float32)
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_model():
    import torch
    from oracle.torch_cpu import TorchCpuModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return TorchCpuModel(model_path()), torch.get_num_threads()


CPU_NOTE = ("torch-CPU fp32 restatement of the Keras graph (oracle/torch_cpu.py) - the reference's TensorFlow-CPU "
            "model.predict is not installable in this image (no tensorflow/keras/h5py wheel, no network); the "
            "reference README quotes ~15 reads/s on 12 threads at 12-24 windows per read")


def cpu_baseline(seconds, x_sample):
    """torch-CPU fp32 oracle on all host cores, batch 256, for ~`seconds` of CPU work."""
    m, cores = cpu_model()
    m.predict(x_sample[:BATCH])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        m.predict(x_sample[(n // BATCH * BATCH) % (len(x_sample) - BATCH + 1):][:BATCH])
        n += BATCH
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': 'reads/s', 'cores': cores, 'kind': 'port',
            'sample': '{} windows (batch {}) of the same synthetic workload in {:.1f} s; {}'.format(n, BATCH, dt, CPU_NOTE)}


def workload_text(shard):
    return ('synthetic 1M-read config (per-GPU shard of {} reads x 1024 samples), EXP-NBD103 start model, '
            'scan_size 512 => 1 window per read, batch {}'.format(shard, BATCH))


def run_reference(args, rank, world_size):
    """--impl reference: the CPU implementation of the path on the box's host cores (oracle port,
    see cpu_baseline) on the same config/metric: a step = one pass over the same 65 536-read shard in
    batches of 256 (the number of steps is capped so that the run ends within a few minutes).
    Rank 0 only."""
    if rank != 0:
        return
    shard = (args.shard // BATCH) * BATCH
    x = synthetic_windows(shard, seed=1000)
    m, cores = cpu_model()
    t0 = time.perf_counter()
    for _ in range(max(args.warmup, 1)):
        m.predict(x[:4 * BATCH], batch_size=BATCH)
    per_read = (time.perf_counter() - t0) / (max(args.warmup, 1) * 4 * BATCH)
    steps = max(1, min(args.steps, int(150.0 / max(per_read * shard, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        m.predict(x, batch_size=BATCH)
    dt = time.perf_counter() - t0
    value = steps * shard / dt
    line = {
        'impl': 'reference', 'metric': 'reads classified/sec', 'value': value, 'unit': 'reads/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_text(shard), 'batch': BATCH, 'reads_per_step_per_gpu': shard,
                   'windows_per_read': 1,
                   'note': 'CPU arm: one host, all cores, the whole shard per step; steps capped at {} of the {} '
                           'requested to keep the run within a few minutes'.format(steps, args.steps)},
        'cpu_baseline': {'value': value, 'unit': 'reads/s', 'cores': cores, 'kind': 'port',
                         'sample': '{} steps x {} reads; {}'.format(steps, shard, CPU_NOTE)},
        'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def unsaturated_parity(model):
    """The parity statistic of the metric on the committed UNSATURATED population (oracle top-1 < 0.99;
    tests/golden/unsaturated_windows.npz, made by tests/golden/make_unsaturated.py): softmax max-abs-err
    against the stored fp64 oracle rows."""
    from oracle import deepbinner_oracle as orc
    u = np.load(ROOT / 'tests' / 'golden' / 'unsaturated_windows.npz')
    z = np.load(ROOT / 'tests' / 'golden' / 'fixture_reads.npz')
    reads = [str(r) for r in u['reads']]
    x = np.stack([orc.normalise(z[reads[r]][o:o + 1024]) for r, o in zip(u[MODEL + '|read'], u[MODEL + '|offset'])])
    ref = u[MODEL + '|probs']
    got = model.predict(x[:, :, None])
    err = np.abs(got - ref).max(axis=1)
    top = ref.max(axis=1)
    return {'max_abs_err': float(err.max()), 'p99': float(np.percentile(err, 99)), 'n': int(len(x)),
            'unsaturated': int((top < 0.99).sum()), 'top1_in_0.3_0.7': int(((top >= 0.3) & (top <= 0.7)).sum()),
            'argmax_mismatches': int((got.argmax(axis=1) != ref.argmax(axis=1)).sum()),
            'vs': 'fp64 oracle (oracle/deepbinner_oracle.py) rows committed in tests/golden/unsaturated_windows.npz',
            'tolerance': 1e-3}


def ragged_reads(n, seed):
    rng = np.random.RandomState(seed)
    return [np.clip(rng.normal(rng.uniform(300, 600), rng.uniform(10, 500), rng.randint(3000, 20000)),
                    -32768, 32767).astype(np.int16) for _ in range(n)]


def pipeline_rate(cls, batches, start, end, args_ns, n_classes, reps):
    """reads/s of classify.classify_read_batches over `reps` passes of `batches` (list of (ids, signals))."""
    def source():
        for _ in range(reps):
            for ids, sigs in batches:
                yield ids, sigs, None
    for _ in range(2):                                                            # warm-up passes (buffers of all job slots)
        cls.classify_read_batches(((i, s, None) for i, s in batches), start, 1024 if start else None, end,
                                  1024 if end else None, n_classes, args_ns)
    rates = []
    for _ in range(3):                                                            # median of three timed runs (each ~0.1 s)
        t0 = time.perf_counter()
        cls.classify_read_batches(source(), start, 1024 if start else None, end, 1024 if end else None,
                                  n_classes, args_ns)
        rates.append(reps * sum(len(i) for i, _ in batches) / (time.perf_counter() - t0))
    return sorted(rates)[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--engine', default=None, choices=[None, 'fp32', 'tcgen05'])
    ap.add_argument('--shard', type=int, default=SHARD)
    ap.add_argument('--config', default='headline', choices=['headline', 'strong'],
                    help="'strong': 1 048 576 reads in total, split over the ranks (strong scaling)")
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the configs[1]/[2]/[4] lines')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    world_size = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world_size)
        return

    import types
    import torch
    from deepbinner_b200 import classify as cls
    from deepbinner_b200 import parallel, weights
    from deepbinner_b200.model import B200Model

    rank, local_rank, world_size = parallel.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = parallel.bind_to_numa_node_of_gpu(local_rank)

    # the one collective of the path: broadcast the packed weights from rank 0 (NCCL)
    blob = weights.load_blob(model_path()) if rank == 0 else None
    blob = parallel.broadcast_blob(blob, dev)
    model = B200Model(blob=blob, device=local_rank, engine=args.engine)

    if args.config == 'strong':
        args.shard = 1048576 // world_size
    shard = (args.shard // BATCH) * BATCH
    n_batches = shard // BATCH
    reads_i16 = synthetic_reads(shard, seed=1000 + rank)
    x_host = torch.empty((shard, 1024), dtype=torch.float32).pin_memory()
    xr = reads_i16.astype(np.float64)
    mu, sg = xr.mean(axis=1, keepdims=True), xr.std(axis=1, keepdims=True)
    sg[sg == 0] = 1.0
    x_host.numpy()[:] = ((xr - mu) / sg).astype(np.float32)     # the same reads, z-scored (seam b1 input)
    del xr
    d_x = x_host.to(dev)
    d_p = torch.zeros((shard, model.n_classes), dtype=torch.float32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(N_STREAMS)]
    main_stream = torch.cuda.current_stream(dev)

    def device_step():
        for b in range(n_batches):
            st = streams[b % N_STREAMS]
            model.predict_device(d_x.data_ptr() + b * BATCH * 4096, BATCH,
                                 d_p.data_ptr() + b * BATCH * model.n_classes * 4, st.cuda_stream)

    def timed_device(steps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        parallel.barrier()
        start.record(main_stream)
        for st in streams:
            st.wait_event(start)
        for _ in range(steps):
            device_step()
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main_stream.wait_event(ev)
        end.record(main_stream)
        torch.cuda.synchronize(dev)
        parallel.barrier()
        return start.elapsed_time(end)

    # ---- kernel-only throughput (inputs resident in HBM) ----
    timed_device(args.warmup)
    launches0 = model.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed_device(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = model.kernel_launches - launches0
    ms_max = parallel.max_over_ranks(ms)
    reads_total = world_size * shard * args.steps
    value = reads_total / (ms_max * 1e-3)

    # ---- large-batch variant (one launch per step over the whole shard), informational ----
    def big_step():
        model.predict_device(d_x.data_ptr(), shard, d_p.data_ptr(), main_stream.cuda_stream)
    big_step()
    torch.cuda.synchronize(dev)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(main_stream)
    for _ in range(max(args.steps // 4, 2)):
        big_step()
    e.record(main_stream)
    torch.cuda.synchronize(dev)
    big_ms = s.elapsed_time(e) / max(args.steps // 4, 2)

    # ---- kernel-only rate of the fused call_batch kernel (int16 reads resident in HBM, z-score on device;
    #      the same launch structure as `value`: batch 256 on 4 streams), informational ----
    d_reads = torch.from_numpy(reads_i16).to(dev)
    d_off = torch.arange(BATCH + 1, dtype=torch.int64, device=dev) * 1024
    d_calls = torch.zeros(shard, dtype=torch.int8, device=dev)
    d_step = torch.zeros((N_STREAMS, BATCH, model.n_classes), dtype=torch.float32, device=dev)

    def call_device_step():
        for b in range(n_batches):
            k = b % N_STREAMS
            model.call_batch_device(d_reads.data_ptr() + b * BATCH * 2048, d_off.data_ptr(), BATCH, 'start', 512, 0.5,
                                    d_p.data_ptr() + b * BATCH * model.n_classes * 4, d_calls.data_ptr() + b * BATCH,
                                    d_step[k].data_ptr(), streams[k].cuda_stream)
    call_device_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        call_device_step()
    torch.cuda.synchronize(dev)
    call_kernel_rate = 5 * shard / (time.perf_counter() - t0)
    del d_reads

    # ---- end to end, headline: the fused call_batch entry on raw int16 reads in host memory ----
    E2E_BATCH = 8192
    reads_pinned = torch.from_numpy(reads_i16).pin_memory().numpy()     # the step's inputs live in pinned host memory
    packed = [PackedReads(reads_pinned[a:a + E2E_BATCH]) for a in range(0, shard, E2E_BATCH)]

    def e2e_steps_run(steps):
        """`steps` passes over the shard: every job's inputs go host -> device and its calls + probabilities come
        back to the host inside this function; up to three jobs are in flight, also across step boundaries."""
        jobs, out = [], None
        for _ in range(steps):
            for pk in packed:
                jobs.append(model.call_batch_async(pk, 'start', 512, 0.5))
                if len(jobs) == 3:
                    out = jobs.pop(0).result()
        for j in jobs:
            out = j.result()
        return out
    e2e_steps_run(1)
    torch.cuda.synchronize(dev)
    parallel.barrier()
    e2e_steps = max(min(args.steps, 10), 3)
    t0 = time.perf_counter()
    last_calls, last_probs = e2e_steps_run(e2e_steps)
    torch.cuda.synchronize(dev)
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0)
    parallel.barrier()
    e2e_value = world_size * shard * e2e_steps / e2e_s

    # ---- end to end through seam b1 (float32 windows, pinned host) ----
    xh = x_host.numpy()
    model.predict(xh, batch_size=BATCH)
    torch.cuda.synchronize(dev)
    parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.predict(xh, batch_size=BATCH)
    e2e_f32_s = parallel.max_over_ranks(time.perf_counter() - t0)
    parallel.barrier()
    e2e_f32 = world_size * shard * e2e_steps / e2e_f32_s

    # ---- BASELINE configs[4] stand-in on EVERY rank: `realtime` without the file system - a streaming source
    #      of already-parsed reads, rounds of 20 480 reads per GPU in batches of 4 096 packed reads, start + end
    #      models, scan 6144, through the product's batch pipeline; whole-job rate = reads of all ranks / slowest rank
    stream_cfg = None
    if not args.no_configs:
        end_all = B200Model(str(ROOT / 'deepbinner_b200' / 'models' / 'EXP-NBD103_read_ends.dbnw'),
                            device=local_rank, engine=args.engine)
        ns5 = types.SimpleNamespace(scan_size=6144.0, batch_size=4096, score_diff=0.5, require_either=True,
                                    require_start=False, require_both=False, verbose=False)
        pool = torch.from_numpy(synthetic_reads(4096, seed=99 + rank, length=2 * (6144 + 512))).pin_memory().numpy()
        pk = PackedReads(pool)
        stream_batches = [(['s%d_%d' % (k, i) for i in range(4096)], pk) for k in range(5)]
        pipeline_rate(cls, stream_batches, model, end_all, ns5, model.n_classes, reps=1)         # warm-up (all job slots)
        torch.cuda.synchronize(dev)
        parallel.barrier()
        t0 = time.perf_counter()
        rounds = 3
        for _ in range(rounds):
            cls.classify_read_batches(((i, sg, None) for i, sg in stream_batches), model, 1024, end_all, 1024,
                                      model.n_classes, ns5)
        stream_s = parallel.max_over_ranks(time.perf_counter() - t0)
        parallel.barrier()
        r5 = world_size * rounds * 5 * 4096 / stream_s
        stream_cfg = {
            'reads_per_s': r5, 'windows_per_s': 24 * r5, 'windows_per_read': 24, 'n_gpus': world_size,
            'h2d_bytes_per_read': 2 * 2 * (6144 + 512), 'rounds': rounds, 'reads_per_round_per_gpu': 5 * 4096,
            'what': 'BASELINE configs[4] stand-in: classify.classify_read_batches (the batch pipeline of `realtime`) over a '
                    'streaming source of packed int16 reads (13 312 samples each, pinned host memory) on every GPU - reads '
                    'are sharded, no collective; rounds of 20 480 reads per GPU in batches of 4 096, start + end models, '
                    'scan 6144; whole-job rate = reads of all ranks / time of the slowest rank'}
        end_all.close()

    configs, n256 = None, None
    if rank == 0 and not args.no_configs:
        # ---- the shape the reference's call_batch really produces at seam b1: predict() with n = 256
        #      pageable float64 windows per call (classify.py:340-361) ----
        x64 = np.ascontiguousarray(xh[:256].astype(np.float64)[:, :, None])
        model.predict(x64, batch_size=256)
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            model.predict(x64, batch_size=256)
        n256 = reps * 256 / (time.perf_counter() - t0)
        # ---- BASELINE configs[1], configs[2], configs[4] through the product's batch pipeline ----
        try:
            configs = {}
            end_model = B200Model(str(ROOT / 'deepbinner_b200' / 'models' / 'EXP-NBD103_read_ends.dbnw'),
                                  device=local_rank, engine=args.engine)
            rapid = B200Model(str(ROOT / 'deepbinner_b200' / 'models' / 'SQK-RBK004_read_starts.dbnw'),
                              device=local_rank, engine=args.engine)
            ns = types.SimpleNamespace(scan_size=6144.0, batch_size=BATCH, score_diff=0.5, require_either=True,
                                       require_start=False, require_both=False, verbose=False)
            ids256 = ['r%d' % i for i in range(256)]
            b256 = [(ids256, ragged_reads(256, 7 + k)) for k in range(4)]
            r = pipeline_rate(cls, b256, model, end_model, ns, model.n_classes, reps=30)
            configs['native_start_end_batch256'] = {
                'reads_per_s': r, 'windows_per_s': 24 * r, 'windows_per_read': 24,
                'what': 'BASELINE configs[1]: classify.classify_read_batches (both sides submitted per batch, '
                        'batches software-pipelined) on host lists of 256 ragged int16 reads, scan 6144'}
            ids512 = ['r%d' % i for i in range(512)]
            b512 = [(ids512, ragged_reads(512, 11 + k)) for k in range(4)]
            ns3 = types.SimpleNamespace(**dict(vars(ns), batch_size=512))
            r = pipeline_rate(cls, b512, rapid, None, ns3, rapid.n_classes, reps=30)
            configs['rapid_start_batch512'] = {
                'reads_per_s': r, 'windows_per_s': 12 * r, 'windows_per_read': 12,
                'what': 'BASELINE configs[2]: SQK-RBK004_read_starts, host lists of 512 ragged int16 reads, scan 6144'}
            end_model.close()
            rapid.close()
            configs['realtime_stream_start_end'] = stream_cfg
        except Exception as ex:  # noqa: BLE001
            configs = {'error': repr(ex)}

    # ---- parity statistic of the metric + CPU baseline ----
    parity, cpu = None, None
    if rank == 0:
        parity = unsaturated_parity(model)
        if not args.no_cpu_baseline and world_size == 1:
            cpu = cpu_baseline(args.cpu_seconds, xh[:8192])

    if rank == 0:
        burst, sustained, peak_src = measured_peaks()
        avg_launch_ms = ms / max(launches, 1)      # one network pass over a batch = one kernel
        achieved = FLOP_PER_WINDOW * BATCH / (avg_launch_ms * 1e-3) / 1e12
        traffic = None
        tfile = ROOT / 'profiles' / 'traffic_r02.json'
        if tfile.exists():
            traffic = json.loads(tfile.read_text()).get(model.engine)
        terms = 1 if model.engine == 'fp32' else 3
        if configs and 'error' not in configs:
            kernel_wps = value / world_size            # windows/s of the kernel on one GPU (1 window per read)
            for c in configs.values():
                wps = c['windows_per_s'] / c.get('n_gpus', 1)
                c['fraction_of_kernel_rate'] = wps / kernel_wps
                c['roofline_frac'] = wps * FLOP_PER_WINDOW / 1e12 / burst
                if cpu is not None:
                    c['cpu_port_reads_per_s'] = cpu['value'] / c['windows_per_read']
        line = {
            'metric': 'reads classified/sec', 'value': value, 'unit': 'reads/s',
            'n_gpus': world_size, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.config == 'strong' else 'weak',
            'vs_baseline': None, 'dtype': 'f32' if model.engine == 'fp32' else 'bf16x3->f32',
            'data': 'synthetic',
            'config': {
                'workload': workload_text(shard), 'batch': BATCH, 'reads_per_step_per_gpu': shard,
                'windows_per_read': 1, 'engine': model.engine, 'streams': N_STREAMS,
                'l2': 'inputs per step ({} MiB fp32) exceed the 126 MB L2; no flush needed'.format(shard * 4096 >> 20),
                'large_batch_reads_per_s': world_size * shard / (big_ms * 1e-3),
                'call_batch_kernel_reads_per_s_per_gpu': call_kernel_rate,
                'numa_node_of_rank0': numa,
                'configs': configs},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'reads/s', 'h2d_bytes_per_step': shard * 1024 * 2,
                    'd2h_bytes_per_step': shard * (model.n_classes * 4 + 1),
                    'api': 'B200Model.call_batch_async(packed int16 reads, side=start, scan_size=512) in jobs of '
                           '{} reads, 3 in flight -> db_call_batch_submit_packed / db_call_batch_wait'.format(E2E_BATCH),
                    'predict_f32': {'value': e2e_f32, 'h2d_bytes_per_step': shard * 4096,
                                    'd2h_bytes_per_step': shard * model.n_classes * 4,
                                    'api': 'B200Model.predict(x_pinned_host[{},1024] float32) -> db_predict_windows'.format(shard)},
                    'predict_n256_f64': n256},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': burst, 'unit': 'TFLOP/s',
                         'frac': achieved / burst, 'frac_sustained': achieved / sustained,
                         'peak_sustained': sustained, 'traffic': traffic,
                         # split-bf16: three tensor-core MMAs per algorithmic one
                         'issued': achieved * terms, 'issued_frac': achieved * terms / burst,
                         'note': 'algorithmic 33,629,952 FLOP/window x {} windows per launch / average launch '
                                 'duration (timed region / launches; launches on {} streams overlap); peak = {}'
                                 .format(BATCH, N_STREAMS, peak_src)},
            'parity': parity,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line), flush=True)
    parallel.barrier()
    if world_size > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
