"""
Command line of the B200 classifier: the `classify` and `realtime` commands of reference
`deepbinner/deepbinner.py` with the same options, presets, defaults and validation messages
(`classify_subparser` :90-106, `classify_and_realtime_options` :109-156, `realtime_subparser`
:177-196, `check_classify_and_realtime_arguments` :283-317, `find_model` :320-345).  The reference's
`bin` (host-side text streaming, bin.py) completes the classify -> bin workflow; the other
sub-commands (prep, balance, train, refine) are outside the accelerated hot path.
The four TensorFlow thread knobs are accepted and ignored; `--device` selects the GPU and `--gpus N`
runs `classify` / `realtime` data-parallel on N GPUs of the box: the command re-launches itself as one
process per GPU (torch.distributed.run, NCCL), rank 0 broadcasts the weights and the file list, every
rank classifies a contiguous shard of the reads, rank 0 prints the rows in rank (= input) order.
"""

import argparse
import os
import pathlib
import sys

from . import __version__

MODEL_DIR = pathlib.Path(__file__).resolve().parent / 'models'


def main(argv=None):
    parser = argparse.ArgumentParser(
        prog='deepbinner',
        description='Deepbinner (B200 engine): a deep convolutional neural network barcode '
                    'demultiplexer for Oxford Nanopore reads')
    parser.add_argument('--version', action='version', version=__version__)
    subparsers = parser.add_subparsers(title='Commands', dest='subparser_name')
    classify_subparser(subparsers)
    realtime_subparser(subparsers)
    bin_subparser(subparsers)

    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        parser.print_help(file=sys.stderr)
        sys.exit(1)
    args = parser.parse_args(argv)

    if getattr(args, 'gpus', 1) > 1 and int(os.environ.get('WORLD_SIZE', '1')) == 1:
        sys.exit(relaunch_on_gpus(args.gpus, argv))

    if args.subparser_name == 'classify':
        check_classify_and_realtime_arguments(args)
        from .classify import classify
        classify(args)
    elif args.subparser_name == 'realtime':
        check_classify_and_realtime_arguments(args)
        from .realtime import realtime
        realtime(args)
    elif args.subparser_name == 'bin':
        from .bin import bin_reads
        bin_reads(args)
    else:
        parser.print_help(file=sys.stderr)
        sys.exit(1)


def relaunch_on_gpus(n_gpus, argv):
    """`--gpus N`: run this very command as N ranks (one process per GPU) under torch.distributed.run;
    returns the exit code."""
    import socket
    import subprocess
    with socket.socket() as sock:
        sock.bind(('127.0.0.1', 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n_gpus),
           '--master-addr', '127.0.0.1', '--master-port', str(port), '-m', 'deepbinner_b200'] + list(argv)
    return subprocess.call(cmd)


def classify_subparser(subparsers):
    group = subparsers.add_parser('classify', description='Classify fast5 reads')
    group.add_argument('input', type=str,
                       help='One of the following: a single fast5 file, a directory of fast5 files '
                            '(will be searched recursively) or a tab-delimited file of training '
                            'data')
    classify_and_realtime_options(group)
    group.add_argument('--verbose', action='store_true',
                       help='Include the output probabilities for all barcodes in the results '
                            '(default: just show the final barcode call)')


def realtime_subparser(subparsers):
    group = subparsers.add_parser('realtime', description='Sort fast5 files during sequencing')
    group.add_argument('--in_dir', type=str, required=True,
                       help='Directory where sequencer deposits fast5 files')
    group.add_argument('--out_dir', type=str, required=True,
                       help='Directory to output binned fast5 files')
    classify_and_realtime_options(group)
    group.add_argument('--stop', action='store_true',
                       help='Automatically stop when there are no more input reads (default: '
                            'continue to run and wait for more reads)')


def bin_subparser(subparsers):
    """`deepbinner bin` (reference deepbinner.py:159-174)."""
    group = subparsers.add_parser('bin', description='Bin fasta/q reads')
    required = group.add_argument_group('Required')
    required.add_argument('--classes', type=str, required=True,
                          help='Deepbinner classification file (made with the deepbinner classify '
                               'command)')
    required.add_argument('--reads', type=str, required=True, help='FASTA or FASTQ reads')
    required.add_argument('--out_dir', type=str, required=True,
                          help='Directory to output binned read files')


def classify_and_realtime_options(group):
    presets = group.add_argument_group('Model presets')
    presets.add_argument('--native', action='store_true',
                         help='Preset for EXP-NBD103 read start and end models')
    presets.add_argument('--rapid', action='store_true',
                         help='Preset for SQK-RBK004 read start model')

    models = group.add_argument_group('Models (at least one is required if not using a preset)')
    models.add_argument('-s', '--start_model', type=str, help='Model trained on the starts of reads')
    models.add_argument('-e', '--end_model', type=str, help='Model trained on the ends of reads')

    barcoding = group.add_argument_group('Barcoding')
    barcoding.add_argument('--scan_size', type=float, default=6144,
                           help="This much of a read's start/end signal will examined for barcode "
                                "signals")
    barcoding.add_argument('--score_diff', type=float, default=0.5,
                           help='For a read to be classified, there must be this much difference '
                                'between the best and second-best barcode scores')

    two = group.add_argument_group('Two model (read start and read end) behaviour')
    two.add_argument('--require_either', action='store_true',
                     help='Most lenient approach: a barcode call on either the start or end is '
                          'sufficient to classify a read, as long as they do not disagree on the '
                          'barcode (default behaviour)')
    two.add_argument('--require_start', action='store_true',
                     help='Moderate approach: a start barcode is required to classify a read but '
                          'an end barcode is optional')
    two.add_argument('--require_both', action='store_true',
                     help='Most stringent approach: both start and end barcodes must be present '
                          'and agree to classify a read')

    perf = group.add_argument_group('Performance')
    perf.add_argument('--batch_size', type=int, default=256, help='Neural network batch size')
    perf.add_argument('--device', type=int, default=0, help='CUDA device ordinal')
    perf.add_argument('--gpus', type=int, default=1,
                      help='Number of GPUs of this machine to spread the reads over (one process per GPU)')
    for knob, default in (('intra_op_parallelism_threads', 12), ('inter_op_parallelism_threads', 1),
                          ('device_count', 1), ('omp_num_threads', 12)):
        perf.add_argument('--' + knob, type=int, default=default,
                          help='Accepted for compatibility with the TensorFlow build; ignored')


def check_classify_and_realtime_arguments(args):
    if args.native and args.rapid:
        sys.exit('Error: you can only use one model preset (--native or --rapid)')
    if args.native or args.rapid:
        if args.start_model is not None or args.end_model is not None:
            sys.exit('Error: you cannot explicitly specify a model and also use a model preset '
                     '(--{})'.format('native' if args.native else 'rapid'))
    if args.native:
        args.start_model = find_native_start_model()
        args.end_model = find_native_end_model()
    if args.rapid:
        args.start_model = find_rapid_start_model()

    model_count = sum(1 for m in (args.start_model, args.end_model) if m is not None)
    if model_count == 0:
        sys.exit('Error: you must provide at least one model')
    if not 0.0 < args.score_diff <= 1.0:
        sys.exit('Error: --score_diff must be in the range (0, 1] (greater than 0 and less than or '
                 'equal to 1)')
    for flag in ('require_either', 'require_start', 'require_both'):
        if model_count < 2 and getattr(args, flag):
            sys.exit('Error: --{} can only be used with two models (start and end)'.format(flag))
    used = two_model_args_used(args)
    if used > 1:
        sys.exit('Error: only one of the following options can be used: --require_either, '
                 '--require_start, --require_both')
    if used == 0:
        args.require_either = True      # the code default of the reference (deepbinner.py:315-316)


def find_native_start_model():
    return find_model('EXP-NBD103_read_starts')


def find_native_end_model():
    return find_model('EXP-NBD103_read_ends')


def find_rapid_start_model():
    return find_model('SQK-RBK004_read_starts')


def find_model(model_name):
    """Locate a shipped model: the packed DBNW blob in this package, else a Keras HDF5 file of that
    name in a `models/` directory next to the package."""
    for candidate in (MODEL_DIR / (model_name + '.dbnw'), MODEL_DIR / model_name,
                      MODEL_DIR.parents[1] / 'models' / model_name):
        if candidate.is_file():
            return str(candidate)
    sys.exit('Error: could not find {} - did Deepbinner install correctly?'.format(model_name))


def two_model_args_used(args):
    return sum(1 for f in (args.require_either, args.require_start, args.require_both) if f)


if __name__ == '__main__':
    main()
