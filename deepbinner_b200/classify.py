"""
`deepbinner classify` on the B200 engine: the host side of the hot path, mirroring the operator
interface of reference `deepbinner/classify.py` (same function names, argument meaning, return
values, stdout/stderr split, TSV format and `sys.exit('Error: ...')` behaviour) so that callers of
`classify_fast5_files` / `call_batch` / `load_and_check_models` can switch over unchanged.

What differs is only where the arithmetic runs: `load_trained_model` returns a `B200Model`
(model.py) instead of a Keras model, and `call_batch` hands the whole batch - windowing,
normalisation, CNN, step merge, renormalisation and call (reference classify.py:325-393,
:285-295) - to one fused GPU call when the model is a B200Model.  A foreign model object exposing
`.predict` (seam b1) still works through the generic host loop.
"""

import collections
import concurrent.futures
import os
import pathlib
import sys

import numpy as np

from . import hdf5_lite, weights
from .load_fast5s import (find_all_fast5s, get_read_id_and_signal, determine_single_or_multi_fast5s,
                          read_fast5_batch, read_fast5_batch_packed)
from .misc import print_summary_table
from .model import B200Model, ReadPointers, signals_fit_int16
from .trim_signal import normalise


def world():
    """(rank, local_rank, world_size) of this process (torchrun environment; 1 process if unset)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


def classify_distributed(args):
    """`classify --gpus N` inside one of the N ranks (one process per GPU): rank 0 reads the model files
    and lists the input, the weights and the file list are broadcast (the one collective of the data
    path), every rank classifies a contiguous shard of the files on its own GPU, and rank 0
    prints header, rows in rank order (= the order of a one-GPU run) and the summary table."""
    import contextlib
    import io
    from . import parallel
    rank, local_rank, world_size = parallel.init()
    device = parallel.device_for(local_rank)
    parallel.bind_to_numa_node_of_gpu(device)
    quiet = open(os.devnull, 'w')
    log = sys.stderr if rank == 0 else quiet
    set_tensorflow_threads(args)
    start_model, start_input_size, end_model, end_input_size, output_size, model_count = \
        load_and_check_models(args.start_model, args.end_model, args.scan_size, out_dest=log, device=device,
                              distributed=True)
    input_type = parallel.broadcast_object(determine_input_type(args.input) if rank == 0 else None)
    if input_type == 'training_data':      # a text file read sequentially: not sharded, rank 0 does it
        if rank == 0:
            if model_count == 2:
                sys.exit('Error: training data can only be classified using a single model')
            print('', file=sys.stderr)
            classify_training_data(args.input, start_model, start_input_size, end_model, end_input_size,
                                   output_size, args)
        parallel.barrier()
        return
    print('', file=log)
    files, multi = None, None
    if rank == 0:
        files = find_all_fast5s(args.input, verbose=True) if input_type == 'directory' else [args.input]
        if not files:
            files = None
        else:
            multi = determine_single_or_multi_fast5s(files)
    files, multi = parallel.broadcast_object((files, multi))
    if files is None:
        sys.exit('Error: no fast5 files found')
    lo, hi = parallel.shard_range(len(files), rank, world_size)
    rows, calls = io.StringIO(), {}
    if hi > lo:
        with contextlib.redirect_stdout(rows), contextlib.redirect_stderr(log):
            calls, _ = classify_fast5_files(files[lo:hi], start_model, start_input_size, end_model, end_input_size,
                                            output_size, args, summary_table=False, print_header=False,
                                            known_layout=multi)
    parts = parallel.gather_objects((rows.getvalue(), calls))
    if rank == 0:
        print_output_header(args.verbose, start_model is not None, end_model is not None, output_size)
        merged = {}
        for text, part in parts:
            sys.stdout.write(text)
            merged.update(part)
        sys.stdout.flush()
        print('', file=sys.stderr)
        print_summary_table(merged)
    parallel.barrier()


def classify(args):
    """Entry point of the `classify` command (reference classify.py:32-55)."""
    if world()[2] > 1:
        return classify_distributed(args)
    set_tensorflow_threads(args)
    start_model, start_input_size, end_model, end_input_size, output_size, model_count = \
        load_and_check_models(args.start_model, args.end_model, args.scan_size,
                              device=getattr(args, 'device', 0))
    input_type = determine_input_type(args.input)
    if input_type == 'training_data' and model_count == 2:
        sys.exit('Error: training data can only be classified using a single model')
    print('', file=sys.stderr)

    if input_type == 'training_data':
        classify_training_data(args.input, start_model, start_input_size, end_model,
                               end_input_size, output_size, args)
        return
    if input_type == 'directory':
        fast5_files = find_all_fast5s(args.input, verbose=True)
    else:
        fast5_files = [args.input]
    classify_fast5_files(fast5_files, start_model, start_input_size, end_model, end_input_size,
                         output_size, args)


def load_and_check_models(start_model_filename, end_model_filename, scan_size, out_dest=sys.stderr,
                          device=0, distributed=False):
    """-> (start_model, start_input_size, end_model, end_input_size, output_size, model_count),
    as reference classify.py:58-83.  `distributed`: rank 0 reads the files, the packed weights are
    broadcast to the other ranks."""
    loaded = {}
    for side, filename in (('start', start_model_filename), ('end', end_model_filename)):
        if filename is None:
            loaded[side] = (None, None, None)
            continue
        model, input_size, output_size = load_trained_model(filename, out_dest=out_dest,
                                                            device=device, distributed=distributed)
        check_input_size(input_size, scan_size)
        loaded[side] = (model, input_size, output_size)
    start_model, start_input_size, start_output_size = loaded['start']
    end_model, end_input_size, end_output_size = loaded['end']
    model_count = sum(1 for m in (start_model, end_model) if m is not None)
    if model_count == 2 and start_output_size != end_output_size:
        sys.exit('Error: two models have different number of barcode classes')
    output_size = start_output_size if start_model is not None else end_output_size
    return start_model, start_input_size, end_model, end_input_size, output_size, model_count


def load_trained_model(model_file, out_dest=sys.stderr, device=0, distributed=False):
    """Load a Keras HDF5 model file (or a DBNW blob) onto the GPU -> (model, input_size,
    output_size) (reference classify.py:86-103)."""
    if distributed and world()[2] > 1:
        return load_trained_model_broadcast(model_file, out_dest, device)
    if not pathlib.Path(model_file).is_file():
        sys.exit('Error: {} does not exist'.format(model_file))
    print('Loading {}... '.format(model_file), file=out_dest, end='', flush=True)
    # (B200 engines: input size 1024 - what every shipped model has - and at most 32 classes; the
    # native layer rejects anything else with DBN_EFORMAT, reported like any other invalid model file.
    # The reference accepts any even input size, check_input_size below.)
    from ._native import NativeError
    try:
        model = B200Model(str(model_file), device=device)
    except (weights.ModelFormatError, hdf5_lite.Hdf5Error, KeyError):
        sys.exit('Error: model input has incorrect shape - are you sure that {} is a valid '
                 'model file?'.format(model_file))
    except NativeError as e:
        sys.exit('Error: could not load {} on the B200 engine: {}'.format(model_file, e))
    print('done', file=out_dest)
    input_size = int(model.inputs[0].shape[1])
    output_size = int(model.outputs[0].shape[1])
    return model, input_size, output_size


def load_trained_model_broadcast(model_file, out_dest, device):
    """Multi-GPU: rank 0 parses the model file; the packed weight blob (0.43 MB) is broadcast to the
    other ranks (NCCL, parallel.broadcast_blob) - the one collective of the data path."""
    from . import parallel
    from ._native import NativeError
    rank = world()[0]
    blob = None
    if rank == 0:
        if not pathlib.Path(model_file).is_file():
            blob = b'!'
        else:
            print('Loading {}... '.format(model_file), file=out_dest, end='', flush=True)
            try:
                blob = weights.load_blob(model_file)
            except (weights.ModelFormatError, hdf5_lite.Hdf5Error, KeyError):
                blob = b'?'
    blob = parallel.broadcast_blob(blob if blob is not None else b'')
    if blob == b'!':
        sys.exit('Error: {} does not exist'.format(model_file))
    if blob == b'?':
        sys.exit('Error: model input has incorrect shape - are you sure that {} is a valid '
                 'model file?'.format(model_file))
    try:
        model = B200Model(blob=blob, device=device)
    except NativeError as e:
        sys.exit('Error: could not load {} on the B200 engine: {}'.format(model_file, e))
    if rank == 0:
        print('done', file=out_dest)
    return model, int(model.inputs[0].shape[1]), int(model.outputs[0].shape[1])


def classify_fast5_files(fast5_files, start_model, start_input_size, end_model, end_input_size,
                         output_size, args, full_output=True, summary_table=True,
                         verified_single_read=False, print_header=True, known_layout=None):
    """Batch driver (reference classify.py:106-180): chunk the file list by `args.batch_size`, load
    signals, call each side, combine, print TSV rows.  -> (classifications, read_id_to_fast5_file)."""
    if not fast5_files:
        sys.exit('Error: no fast5 files found')
    out_dest = sys.stderr if full_output else sys.stdout

    # Multi-read fast5 files are read natively, every read straight out of the file (the reference
    # needs them unpacked to one file per read first, realtime.py:183-196); a batch is then one file
    # (thousands of reads in a MinKNOW file) instead of `batch_size` files.
    # (`known_layout`: 'single' / 'multi' when the caller has already looked - a shard of a multi-GPU run)
    if known_layout is not None:
        multi = known_layout == 'multi'
    else:
        multi = (not verified_single_read) and determine_single_or_multi_fast5s(fast5_files) == 'multi'

    use_start, use_end = start_model is not None, end_model is not None
    print_classification_progress(0, len(fast5_files), 'fast5s', out_dest=out_dest)
    if full_output and print_header:
        print_output_header(args.verbose, use_start, use_end, output_size)

    input_size = start_input_size if use_start else end_input_size
    keep = int(args.scan_size) + input_size // 2
    classifications, read_id_to_fast5_file = {}, {}
    # Each batch is parsed on native host threads (only the samples call_batch can look at - the
    # first / last scan_size + input_size/2 - are kept, which gives identical calls); the parse of
    # batch i+1 runs in the background (the C call releases the GIL) while batch i is on the GPU.
    batches = list(chunker(fast5_files, 1 if multi else args.batch_size))
    prefetcher = concurrent.futures.ThreadPoolExecutor(max_workers=1)
    files_done = [0]

    sides = (1 if use_start else 0) | (2 if use_end else 0)   # which ends of the reads will be looked at

    def parsed_batches():
        pending = prefetcher.submit(load_batch, batches[0], keep, sides)
        for index, batch in enumerate(batches):
            read_ids, signals, kept = pending.result()   # unreadable files are skipped (reference :135-136)
            if index + 1 < len(batches):
                pending = prefetcher.submit(load_batch, batches[index + 1], keep, sides)
            for read_id, file_index in zip(read_ids, kept):
                read_id_to_fast5_file[read_id] = batch[file_index]
            yield read_ids, signals, len(batch)

    def progress(n_files):
        files_done[0] += n_files
        print_classification_progress(files_done[0] if multi else len(classifications), len(fast5_files), 'fast5s',
                                      out_dest=out_dest)

    classify_read_batches(parsed_batches(), start_model, start_input_size, end_model, end_input_size,
                          output_size, args, classifications, print_rows=full_output, progress=progress)

    prefetcher.shutdown()
    if full_output:
        print('', file=sys.stderr)
        if summary_table:
            print_summary_table(classifications)
    return classifications, read_id_to_fast5_file


def classify_read_batches(batches, start_model, start_input_size, end_model, end_input_size, output_size,
                          args, classifications=None, print_rows=False, progress=None):
    """The per-batch part of classify_fast5_files (reference classify.py:141-171) for ANY source of
    reads: `batches` yields (read_ids, signals, tag); each batch is called on both sides, combined and
    (print_rows) printed as TSV rows; `progress(tag)` is invoked after every batch.  -> classifications.
    Software pipeline: the GPU jobs of batch i (start and end side, submitted back to back so that the
    host gathers the end side while the start side computes) are in flight while the results of batch
    i-2 are collected / printed and the source produces batch i+1 (at most three batches, i.e. three of
    a model's four job slots, are pending).  Used by classify_fast5_files (fast5
    parsing as the source), by `realtime`, and with a streaming read source (bench.py, config
    'realtime streaming')."""
    if classifications is None:
        classifications = {}
    use_start, use_end = start_model is not None, end_model is not None

    def finish(read_ids, start_job, end_job, tag):
        """Collect the two sides of a batch, combine, print its TSV rows (reference :150-171)."""
        start_calls, start_probs = start_job() if use_start else (None, None)
        end_calls, end_probs = end_job() if use_end else (None, None)
        for i, read_id in enumerate(read_ids):
            if use_start and use_end:
                final_call = combine_calls(start_calls[i], end_calls[i], args)
            else:
                final_call = start_calls[i] if use_start else end_calls[i]
            classifications[read_id] = final_call
            if not print_rows:
                continue
            row = [read_id, final_call]
            if args.verbose:
                if use_start:
                    row += ['%.2f' % p for p in start_probs[i]]
                    if use_end:
                        row.append(start_calls[i])
                if use_end:
                    row += ['%.2f' % p for p in end_probs[i]]
                    if use_start:
                        row.append(end_calls[i])
            print('\t'.join(row))
        if progress is not None:
            progress(tag)

    in_flight = collections.deque()       # up to two batches behind the one being submitted
    both_b200 = all(isinstance(m, B200Model) for m in (start_model, end_model) if m is not None)
    for read_ids, signals, tag in batches:
        if both_b200 and read_ids and not hasattr(signals, 'samples') and signals_fit_int16(signals):
            signals = ReadPointers(signals)       # pointer / length arrays built once for both sides
        start_job = submit_call_batch(start_input_size, output_size, read_ids, signals, start_model, args,
                                      'start') if use_start else None
        end_job = submit_call_batch(end_input_size, output_size, read_ids, signals, end_model, args,
                                    'end') if use_end else None
        in_flight.append((read_ids, start_job, end_job, tag))
        if len(in_flight) > PIPELINE_DEPTH:
            finish(*in_flight.popleft())
    while in_flight:
        finish(*in_flight.popleft())
    return classifications


def load_batch(fast5_batch, keep, sides=3):
    """(read_ids, signals, kept file indices) of the readable files of a batch, parsed on native host
    threads (reference classify.py:133-139 calls get_read_id_and_signal per file).  `signals` is a
    load_fast5s.PackedSignals: a list of int16 views plus the packed buffer behind them.  `sides`: which
    ends of the reads the models will look at (1 start, 2 end, 3 both)."""
    return read_fast5_batch_packed(fast5_batch, keep=keep, sides=sides)


def classify_training_data(input_file, start_model, start_input_size, end_model, end_input_size,
                           output_size, args):
    """Classify a tab-delimited training file `label<TAB>comma,separated,ints` with one model, rows
    named `line_<n>_barcode_<label>` (reference classify.py:183-239)."""
    use_start, use_end = start_model is not None, end_model is not None
    assert not (use_start and use_end)
    model, input_size = (start_model, start_input_size) if use_start else (end_model, end_input_size)

    with open(input_file, 'rt') as f:
        num_lines = sum(1 for _ in f)
    print_classification_progress(0, num_lines, 'training data')
    print_output_header(args.verbose, use_start, use_end, output_size)

    classifications = {}

    def flush(read_ids, signals):
        calls, probs = call_batch(input_size, output_size, read_ids, signals, model, args, 'start')
        for i, read_id in enumerate(read_ids):
            classifications[read_id] = calls[i]
            row = [read_id, calls[i]]
            if args.verbose:
                row += ['%.2f' % p for p in probs[i]]
            print('\t'.join(row))
        print_classification_progress(len(classifications), num_lines, 'training data')

    read_ids, signals = [], []
    with open(input_file, 'rt') as training_data:
        for line_num, line in enumerate(training_data, start=1):
            barcode, signal = line.rstrip().split('\t')
            read_ids.append('line_{}_barcode_{}'.format(line_num, barcode))
            signals.append(np.array([int(x) for x in signal.split(',')]))
            if len(read_ids) == args.batch_size:
                flush(read_ids, signals)
                read_ids, signals = [], []
    if read_ids or not classifications:
        flush(read_ids, signals)

    print('', file=sys.stderr)
    print_summary_table(classifications)


def determine_input_type(input_file_or_dir):
    """'directory' | 'single_fast5' | 'training_data' (reference classify.py:242-263)."""
    path = pathlib.Path(input_file_or_dir)
    if path.is_dir():
        return 'directory'
    if not path.is_file():
        sys.exit('Error: {} is neither a file nor a directory'.format(input_file_or_dir))
    try:
        hdf5_lite.open_file(input_file_or_dir).close()
        return 'single_fast5'
    except OSError:
        pass
    try:
        with open(input_file_or_dir) as f:
            parts = f.readline().split('\t')
        int(parts[0])
        if len([int(x) for x in parts[1].split(',')]) <= 10:
            raise ValueError
        return 'training_data'
    except (ValueError, IndexError, UnicodeDecodeError):
        sys.exit('Error: could not determine input type')


def chunker(seq, size):
    return (seq[pos:pos + size] for pos in range(0, len(seq), size))


def print_output_header(verbose, using_read_starts, using_read_ends, output_size):
    """TSV header (reference classify.py:270-282; pinned by tests/test_classify.py:198-296)."""
    header = ['read_ID', 'barcode_call']
    barcodes = [str(i) for i in range(1, output_size)]
    if verbose and using_read_starts and using_read_ends:
        for side in ('start', 'end'):
            header += [side + '_none'] + [side + '_' + b for b in barcodes]
            header.append(side + '_barcode_call')
    elif verbose:
        header += ['none'] + barcodes
    print('\t'.join(header))


def get_barcode_call_from_probabilities(probabilities, score_diff_threshold):
    """'none' if class 0 wins or the margin over the runner-up is below the threshold, else the
    winning class as a string; ties resolve to the lower index (reference classify.py:285-295)."""
    probabilities = list(probabilities)
    best = max(range(len(probabilities)), key=lambda j: (probabilities[j], -j))
    if best == 0:
        return 'none'
    runner_up = max(p for j, p in enumerate(probabilities) if j != best)
    return str(best) if probabilities[best] - runner_up >= score_diff_threshold else 'none'


def combine_calls(start_call, end_call, args):
    """Two-model policy table (reference classify.py:298-322; tests/test_combine_calls.py)."""
    if start_call == end_call:
        return start_call
    if args.require_both:
        return 'none'
    if args.require_start:
        return start_call if end_call == 'none' else 'none'
    assert args.require_either
    if start_call == 'none':
        return end_call
    if end_call == 'none':
        return start_call
    return 'none'


def _steps_for(input_size, scan_size):
    step_size = input_size // 2
    steps = int(scan_size / step_size)
    assert steps * step_size == scan_size
    return step_size, steps


_CALL_NAMES = ['none'] + [str(i) for i in range(1, 128)]
PIPELINE_DEPTH = int(os.environ.get('DEEPBINNER_B200_PIPELINE_DEPTH', '2'))   # batches in flight behind the one being submitted


def submit_call_batch(input_size, output_size, read_ids, signals, model, args, side):
    """call_batch in two halves: submits the GPU job of one side of a batch (B200Model.call_batch_async)
    and returns a function that waits for it and returns what call_batch returns.  A foreign model
    (seam b1) is evaluated on the spot."""
    assert side in ('start', 'end')
    _steps_for(input_size, args.scan_size)
    if read_ids and isinstance(model, B200Model) and signals_fit_int16(signals):
        job = model.call_batch_async(signals, side, int(args.scan_size), args.score_diff)

        def collect():
            calls, probs = job.result()
            return [_CALL_NAMES[c] for c in calls.tolist()], list(probs.astype(np.float64))
        return collect
    out = call_batch(input_size, output_size, read_ids, signals, model, args, side)
    return lambda: out


def call_batch(input_size, output_size, read_ids, signals, model, args, side):
    """-> (barcode_calls: list[str], probabilities: list of per-class sequences), index-aligned
    with read_ids (reference classify.py:325-384)."""
    assert side in ('start', 'end')
    _, steps = _steps_for(input_size, args.scan_size)
    if not read_ids:
        return [], []

    if isinstance(model, B200Model) and signals_fit_int16(signals):
        return submit_call_batch(input_size, output_size, read_ids, signals, model, args, side)()

    # Generic path for any object with .predict (seam b1): host windowing, device/foreign predict.
    merged = None
    for s in range(steps):
        windows = build_windows(signals, input_size, s, side)
        labels = np.asarray(model.predict(windows[:, :, np.newaxis], batch_size=args.batch_size))
        if merged is None:
            merged = np.array(labels, dtype=np.float32, copy=True)
        else:
            merged[:, 0] = np.minimum(merged[:, 0], labels[:, 0])
            merged[:, 1:] = np.maximum(merged[:, 1:], labels[:, 1:])
    barcode_calls, probabilities = [], []
    for row in merged:
        p = make_sum_to_one([float(v) for v in row])
        probabilities.append(p)
        barcode_calls.append(get_barcode_call_from_probabilities(p, args.score_diff))
    return barcode_calls, probabilities


def build_windows(signals, input_size, step_index, side):
    """The float64 [n, input_size] network input of scan step `step_index`: slice (mirrored from the
    tail for side 'end'), z-score over the available samples, zero-pad right ('start') or left
    ('end') (reference classify.py:337-358)."""
    lo = step_index * (input_size // 2)
    hi = lo + input_size
    windows = np.zeros((len(signals), input_size), dtype=np.float64)
    for i, signal in enumerate(signals):
        n = len(signal)
        piece = signal[lo:hi] if side == 'start' else signal[max(n - hi, 0):max(n - lo, 0)]
        piece = normalise(piece)
        if len(piece) == 0:
            continue
        if side == 'start':
            windows[i, :len(piece)] = piece
        else:
            windows[i, input_size - len(piece):] = piece
    return windows


def make_sum_to_one(probabilities):
    """Scale the barcode probabilities so that the row sums to one while keeping the no-barcode
    probability (reference classify.py:387-393)."""
    none_prob = probabilities[0]
    scale = (1.0 - none_prob) / sum(probabilities[1:])
    return [none_prob] + [p * scale for p in probabilities[1:]]


def check_input_size(input_size, scan_size):
    """Exit with the reference's messages if the model input size is odd or scan_size is not a
    multiple of half of it (reference classify.py:396-407; tests/test_classify.py:63-68)."""
    step_size = input_size // 2
    if step_size * 2 != input_size:
        sys.exit('Error: the model input size must be even (currently {})'.format(input_size))
    if int(scan_size / step_size) * step_size != scan_size:
        acceptable = [str(step_size * i) for i in range(2, 8)] + ['etc']
        sys.exit('Error: --scan_size must be a multiple of half the model input size\n'
                 'acceptable values for --scan_size are {}'.format(', '.join(acceptable)))


def print_classification_progress(completed, total, label, out_dest=sys.stderr):
    percent = 100.0 * completed / total if total else 100.0
    print('\rClassifying {}: {} / {} ({:.1f}%)'.format(label, completed, total, percent),
          file=out_dest, end='', flush=True)


def set_tensorflow_threads(args):
    """The reference configures TensorFlow's CPU thread pools here (classify.py:416-423).  There is
    no TensorFlow in this path; the four knobs are accepted for command-line compatibility and only
    OMP_NUM_THREADS is exported (host-side numpy/zlib may use it)."""
    omp = getattr(args, 'omp_num_threads', None)
    if omp:
        os.environ['OMP_NUM_THREADS'] = str(omp)
