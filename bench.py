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
e2e    = the same through the public host API `B200Model.predict` (C-ABI db_predict_windows) with
         pinned HOST buffers: H2D of the step's windows and D2H of its probabilities inside the timed
         region.
roofline = tensor (dense conv contraction): algorithmic 33,629,952 FLOP per window.
cpu_baseline = the torch-CPU fp32 oracle on all host cores, bounded sample (the reference's own
         TensorFlow-CPU model.predict cannot run in this image - no tensorflow/keras/h5py).
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
N_STREAMS = 4


def model_path():
    return str(ROOT / 'deepbinner_b200' / 'models' / (MODEL + '.dbnw'))


def measured_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return d.get('bf16_tflops_sustained', d.get('bf16_tflops')), 'MEASURED_PEAKS.json bf16_tflops_sustained'
    return 1400.0, 'fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)'


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


def cpu_baseline(seconds, x_sample):
    """torch-CPU fp32 oracle on all host cores, batch 256, for ~`seconds` of CPU work."""
    import torch
    from oracle.torch_cpu import TorchCpuModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = TorchCpuModel(model_path())
    m.predict(x_sample[:BATCH])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        m.predict(x_sample[(n // BATCH * BATCH) % (len(x_sample) - BATCH + 1):][:BATCH])
        n += BATCH
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': 'reads/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '{} windows (batch {}) of the same synthetic workload in {:.1f} s; torch-CPU fp32 '
                      'restatement of the Keras graph (oracle/torch_cpu.py) - the reference\'s '
                      'TensorFlow-CPU model.predict is not installable here; the reference README '
                      'quotes ~15 reads/s (12 threads, 12-24 windows/read)'.format(n, BATCH, dt)}


def run_reference(args, rank, world_size):
    """--impl reference: the CPU implementation of the path on the box's host cores (oracle port,
    see cpu_baseline) on the same config/metric.  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle.torch_cpu import TorchCpuModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    per_step = 4 * BATCH                      # bounded sample of the workload per step
    x = synthetic_windows(per_step, seed=1234)
    m = TorchCpuModel(model_path())
    for _ in range(max(args.warmup, 1)):
        m.predict(x[:BATCH])
    steps = min(args.steps, 40)
    t0 = time.perf_counter()
    for _ in range(steps):
        m.predict(x, batch_size=BATCH)
    dt = time.perf_counter() - t0
    value = steps * per_step / dt
    line = {
        'impl': 'reference', 'metric': 'reads classified/sec', 'value': value, 'unit': 'reads/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'synthetic 1024-sample reads x EXP-NBD103 start model, scan_size 512 '
                               '(1 window/read), batch 256; CPU sample of {} reads per step'.format(per_step),
                   'batch': BATCH},
        'cpu_baseline': {'value': value, 'unit': 'reads/s', 'cores': torch.get_num_threads(),
                         'kind': 'port',
                         'sample': '{} steps x {} reads; torch-CPU fp32 restatement (oracle/torch_cpu.py); '
                                   'the reference TensorFlow/Keras stack is not installable in this '
                                   'image'.format(steps, per_step)},
        'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--engine', default=None, choices=[None, 'fp32', 'tcgen05'])
    ap.add_argument('--shard', type=int, default=SHARD)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    world_size = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world_size)
        return

    import torch
    from deepbinner_b200 import parallel, weights
    from deepbinner_b200.model import B200Model

    rank, local_rank, world_size = parallel.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)

    # the one collective of the path: broadcast the packed weights from rank 0 (NCCL)
    blob = weights.load_blob(model_path()) if rank == 0 else None
    blob = parallel.broadcast_blob(blob, dev)
    model = B200Model(blob=blob, device=local_rank, engine=args.engine)

    shard = (args.shard // BATCH) * BATCH
    n_batches = shard // BATCH
    x_host = torch.empty((shard, 1024), dtype=torch.float32).pin_memory()
    x_host.numpy()[:] = synthetic_windows(shard, seed=1000 + rank)
    p_host = torch.empty((shard, model.n_classes), dtype=torch.float32).pin_memory()
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

    # ---- end to end through the public host API (pinned host buffers, H2D + D2H inside) ----
    xh = x_host.numpy()
    def e2e_step():
        return model.predict(xh, batch_size=BATCH)
    e2e_step()
    torch.cuda.synchronize(dev)
    parallel.barrier()
    e2e_steps = max(min(args.steps, 10), 3)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        probs_e2e = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0)
    parallel.barrier()
    e2e_value = world_size * shard * e2e_steps / e2e_s

    # ---- informational: BASELINE configs[1] (EXP-NBD103 start+end models, batch 256, default scan
    # 6144 => 12 windows per read per model) through the fused call_batch entry with host buffers ----
    native = None
    if rank == 0:
        try:
            from deepbinner_b200 import classify as cls
            import types
            end_model = B200Model(str(ROOT / 'deepbinner_b200' / 'models' / 'EXP-NBD103_read_ends.dbnw'),
                                  device=local_rank, engine=args.engine)
            rng = np.random.RandomState(7)
            reads = [np.clip(rng.normal(rng.uniform(300, 600), rng.uniform(10, 500), rng.randint(3000, 20000)),
                             -32768, 32767).astype(np.int16) for _ in range(BATCH)]
            ids = ['r%d' % i for i in range(BATCH)]
            a = types.SimpleNamespace(scan_size=6144.0, batch_size=BATCH, score_diff=0.5, require_either=True,
                                      require_start=False, require_both=False)
            def native_step():
                sc, _ = cls.call_batch(1024, model.n_classes, ids, reads, model, a, 'start')
                ec, _ = cls.call_batch(1024, model.n_classes, ids, reads, end_model, a, 'end')
                return [cls.combine_calls(x, y, a) for x, y in zip(sc, ec)]
            native_step()
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                native_step()
            dt = time.perf_counter() - t0
            native = {'reads_per_s': reps * BATCH / dt, 'windows_per_s': reps * BATCH * 24 / dt,
                      'what': 'classify.call_batch x2 + combine_calls on 256 host reads (fused GPU entry, '
                              'H2D/D2H and Python packing included)'}
            end_model.close()
        except Exception as e:  # noqa: BLE001
            native = {'error': str(e)}

    # ---- parity statistic of the metric: softmax max-abs-err vs the CPU reference ----
    parity = None
    cpu = None
    if rank == 0:
        from oracle.torch_cpu import TorchCpuModel
        sample = np.concatenate([xh[:1024:2], xh[::max(shard // 512, 1)][:512]])
        ref = TorchCpuModel(model_path()).predict(sample)
        got = model.predict(sample)
        err = np.abs(got - ref).max(axis=1)
        parity = {'max_abs_err': float(err.max()), 'p99': float(np.percentile(err, 99)),
                  'n': int(len(sample)), 'unsaturated': int((ref.max(axis=1) < 0.99).sum()),
                  'argmax_mismatches': int((got.argmax(axis=1) != ref.argmax(axis=1)).sum()),
                  'vs': 'oracle/torch_cpu.py fp32'}
        if not args.no_cpu_baseline and world_size == 1:
            cpu = cpu_baseline(args.cpu_seconds, xh[:8192])

    if rank == 0:
        peak, peak_src = measured_peaks()
        avg_launch_ms = ms / max(launches, 1)      # one network pass over a batch = one kernel
        achieved = FLOP_PER_WINDOW * BATCH / (avg_launch_ms * 1e-3) / 1e12
        traffic = None
        tfile = ROOT / 'profiles' / 'traffic_r01.json'
        if tfile.exists():
            traffic = json.loads(tfile.read_text()).get(model.engine)
        line = {
            'metric': 'reads classified/sec', 'value': value, 'unit': 'reads/s',
            'n_gpus': world_size, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32' if model.engine == 'fp32' else 'bf16x3->f32',
            'data': 'synthetic',
            'config': {
                'workload': 'synthetic 1M-read config (per-GPU shard of {} reads x 1024 float32 '
                            'samples), EXP-NBD103 start model, scan_size 512 => 1 window per read, '
                            'launches of batch {} over {} streams'.format(shard, BATCH, N_STREAMS),
                'batch': BATCH, 'reads_per_step_per_gpu': shard, 'engine': model.engine,
                'l2': 'inputs per step (256 MiB) exceed L2; no flush needed',
                'large_batch_reads_per_s': shard / (big_ms * 1e-3),
                'native_preset_batch256': native,
                'windows_per_read': 1},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'reads/s', 'h2d_bytes_per_step': shard * 4096,
                    'd2h_bytes_per_step': shard * model.n_classes * 4,
                    'api': 'B200Model.predict(x_pinned_host[{},1024] float32, batch_size=256) -> '
                           'db_predict_windows'.format(shard)},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         # split-bf16: three tensor-core MMAs per algorithmic one
                         'issued': achieved * (1 if model.engine == 'fp32' else 3),
                         'issued_frac': achieved * (1 if model.engine == 'fp32' else 3) / peak,
                         'note': 'algorithmic 33,629,952 FLOP/window x {} windows per launch / '
                                 'average launch duration (timed region / launches; launches on {} '
                                 'streams overlap); peak = {}'.format(BATCH, N_STREAMS, peak_src)},
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
