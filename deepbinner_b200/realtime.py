"""
`deepbinner realtime`: watch a directory the sequencer writes fast5 files into, classify new files
on the GPU and move them into per-barcode directories.  Behaviour follows reference
`deepbinner/realtime.py` (`realtime` :28-70, `look_for_new_fast5s` :73-78, `classify_and_move`
:81-108, `move_classified_fast5s` :111-143, `get_directory_name` :146-150): 5 s polling, at most
20 000 single-read files per round, files whose destination already exists are ignored from then
on, `--stop` ends the loop when nothing is waiting, Ctrl-C exits cleanly.

Multi-read fast5 files.  The reference (`classify_and_move` :93-101, `unpack_multi_read_fast5s`
:183-196) takes at most 5 such files per round, adds them to `ignore_files` (they stay where they
are), unpacks them with ONT's `multi_to_single_fast5` into a temporary directory and bins the
unpacked one-read files.  Here every read is classified straight out of the multi-read file (native
reader, no unpacking, no ONT tool).  The rule matched: same 5-files-per-round limit, the multi-read
file is left in place and ignored from then on, and the per-read result - which the reference
materialises as one-read files under barcodeNN/ - is appended as `read_ID<TAB>barcode_call<TAB>file`
rows to `<out_dir>/multi_read_classifications.tsv` (the first two columns are what `deepbinner bin`
consumes).  Writing one-read fast5 files needs an HDF5 writer and is outside this path.
"""

import os
import pathlib
import shutil
import sys
import time

from .classify import classify_fast5_files, load_and_check_models, set_tensorflow_threads
from .load_fast5s import determine_single_or_multi_fast5s
from .misc import print_summary_table

MAX_FILES_PER_ROUND = 20000
MAX_MULTI_FILES_PER_ROUND = 5
MULTI_TSV = 'multi_read_classifications.tsv'
POLL_SECONDS = 5


def realtime(args, poll_seconds=POLL_SECONDS):
    """`--gpus N` (one process per GPU, see deepbinner.py): rank 0 watches the directory and broadcasts
    each round's file list; every rank classifies and moves a contiguous shard of it on its own GPU
    (moves are independent); the calls are gathered on rank 0 for the summary table."""
    from .classify import world
    rank, local_rank, world_size = world()
    par = None
    device = getattr(args, 'device', 0)
    if world_size > 1:
        from . import parallel as par
        par.init()
        device = par.device_for(local_rank)
        par.bind_to_numa_node_of_gpu(device)
        if rank != 0:
            sys.stdout = open(os.devnull, 'w')     # rank 0 reports; the other ranks only work
    args.verbose = False
    set_tensorflow_threads(args)
    start_model, start_input_size, end_model, end_input_size, output_size, _ = \
        load_and_check_models(args.start_model, args.end_model, args.scan_size,
                              out_dest=sys.stdout, device=device, distributed=world_size > 1)
    in_dir = pathlib.Path(args.in_dir)
    out_dir = pathlib.Path(args.out_dir)
    if rank == 0:
        make_output_dir(out_dir)
    nested_out_dir = is_inside(out_dir, in_dir)

    print('\nLooking for new fast5 files in {}'.format(in_dir), flush=True)
    ignore_files = set()
    waiting_dots = 0
    try:
        while True:
            fast5s, single_or_multi = None, None
            if rank == 0:
                fast5s = [f for f in look_for_new_fast5s(in_dir, out_dir, nested_out_dir)
                          if f not in ignore_files]
                single_or_multi = determine_single_or_multi_fast5s(fast5s) if fast5s else 'single'
            if par:
                fast5s, single_or_multi = par.broadcast_object((fast5s, single_or_multi))
            if fast5s:
                if waiting_dots:
                    print('', flush=True)
                    waiting_dots = 0
                print('\nFound {:,} fast5 files'.format(len(fast5s)), flush=True)
                time.sleep(poll_seconds)   # let files that are still being written settle
                classify_and_move(fast5s, args, start_model, start_input_size, end_model,
                                  end_input_size, output_size, out_dir, ignore_files, single_or_multi, par)
                print('\nLooking for new fast5 files in {}'.format(in_dir), flush=True)
            elif args.stop:
                break
            else:
                print('.', end='', flush=True)
                waiting_dots += 1
                if waiting_dots >= 80:
                    print('', flush=True)
                    waiting_dots = 0
                time.sleep(poll_seconds)
    except KeyboardInterrupt:
        print('\n\nStopping Deepbinner real-time binning', flush=True)


def is_inside(child, parent):
    try:
        child.resolve().relative_to(parent.resolve())
        return True
    except ValueError:
        return False


def look_for_new_fast5s(in_dir, out_dir, nested_out_dir):
    found = [str(p) for p in sorted(pathlib.Path(in_dir).glob('**/*.fast5'))]
    if nested_out_dir:
        prefix = str(pathlib.Path(out_dir).resolve()) + os.sep
        found = [f for f in found if not str(pathlib.Path(f).resolve()).startswith(prefix)]
    return found


def classify_and_move(fast5s, args, start_model, start_input_size, end_model, end_input_size,
                      output_size, out_dir, ignore_files, single_or_multi='single', par=None):
    """One round.  `par` (deepbinner_b200.parallel, multi-GPU runs): this rank takes a contiguous shard
    of the round's files; calls and newly ignored files are gathered afterwards."""
    limit = MAX_FILES_PER_ROUND if single_or_multi == 'single' else MAX_MULTI_FILES_PER_ROUND
    if par:
        limit *= par.world()[2]       # the per-round limit applies per GPU
    if len(fast5s) > limit:       # reference realtime.py:89-94: lots of one-read files, a few multi-read ones
        fast5s = fast5s[:limit]
        print('Limiting this round to {:,} files'.format(limit), flush=True)
    if single_or_multi == 'multi':
        ignore_files.update(fast5s)     # reference :99: the multi-read files themselves stay where they are
    mine = fast5s
    if par:
        rank, _, world_size = par.world()
        lo, hi = par.shard_range(len(fast5s), rank, world_size)
        mine = fast5s[lo:hi]
    classifications, read_id_to_fast5_file = {}, {}
    newly_ignored = set()
    if mine:
        classifications, read_id_to_fast5_file = \
            classify_fast5_files(mine, start_model, start_input_size, end_model, end_input_size,
                                 output_size, args, full_output=False,
                                 verified_single_read=(single_or_multi == 'single'), known_layout=single_or_multi)
        print('', flush=True)
        if single_or_multi == 'multi':
            record_multi_read_classifications(classifications, read_id_to_fast5_file, out_dir,
                                              suffix='.rank{}'.format(par.world()[0]) if par else '')
        else:
            move_classified_fast5s(classifications, read_id_to_fast5_file, out_dir, mine, newly_ignored)
    if par:
        merged = {}
        for calls, ignored in par.all_gather_objects((classifications, newly_ignored)):
            merged.update(calls)
            ignore_files.update(ignored)
        classifications = merged
    else:
        ignore_files.update(newly_ignored)
    print_summary_table(classifications, output=sys.stdout)


def record_multi_read_classifications(classifications, read_id_to_fast5_file, out_dir, suffix=''):
    """Per-read result of multi-read input: rows appended to <out_dir>/multi_read_classifications.tsv
    (multi-GPU runs: one file per rank, `.rank<r>` appended)."""
    path = pathlib.Path(out_dir) / (MULTI_TSV + suffix)
    new = not path.exists()
    with open(str(path), 'at') as f:
        if new:
            f.write('read_ID\tbarcode_call\tfast5_file\n')
        for read_id, call in classifications.items():
            f.write('{}\t{}\t{}\n'.format(read_id, call, read_id_to_fast5_file[read_id]))
    print('Recorded {:,} reads of {:,} multi-read fast5 files in {}'.format(
        len(classifications), len(set(read_id_to_fast5_file.values())), path), flush=True)


def move_classified_fast5s(classifications, read_id_to_fast5_file, out_dir, fast5s, ignore_files):
    """Reference realtime.py:111-143: a file whose destination exists is ignored from then on (and
    counted), a failed move is counted; the watcher only gives up when EVERY file of the round failed
    to move for another reason than an existing destination."""
    move_count, fail_move_already_exists, fail_move_other_reason = 0, 0, 0
    for read_id, barcode_call in classifications.items():
        source = read_id_to_fast5_file[read_id]
        dest_dir = pathlib.Path(out_dir) / get_directory_name(barcode_call)
        if not dest_dir.is_dir():
            try:
                os.makedirs(str(dest_dir))
            except OSError:
                sys.exit('Error: unable to create output directory {}'.format(dest_dir))
        dest = dest_dir / pathlib.Path(source).name
        if dest.is_file():
            fail_move_already_exists += 1
            ignore_files.add(source)
        else:
            try:
                shutil.move(source, str(dest_dir))
                move_count += 1
            except OSError:
                fail_move_other_reason += 1
        print_moving_progress(move_count, len(fast5s))
    print('', flush=True)
    print_moving_error_messages(fail_move_already_exists, fail_move_other_reason, out_dir)
    if fast5s and fail_move_other_reason == len(fast5s):
        sys.exit('Error: no files were successfully moved to {}'.format(out_dir))


def print_moving_progress(completed, total):
    print('\rMoving fast5s:      {} / {} ({:.1f}%)'.format(completed, total, 100.0 * completed / max(total, 1)),
          end='', flush=True)


def print_moving_error_messages(already_exists, other_reason, out_dir):
    if already_exists == 1:
        print('Error: could not move 1 fast5 file because it already exists in {}'.format(out_dir))
    elif already_exists > 1:
        print('Error: could not move {} fast5 files because they already exist '
              'in {}'.format(already_exists, out_dir))
    if other_reason == 1:
        print('Error: failed to move 1 fast5 file to {}'.format(out_dir))
    elif other_reason > 1:
        print('Error: failed to move {} fast5 files to {}'.format(other_reason, out_dir))


def get_directory_name(barcode_call):
    if barcode_call == 'none':
        return 'unclassified'
    return 'barcode{:02d}'.format(int(barcode_call))


def make_output_dir(out_dir):
    out_dir = pathlib.Path(out_dir)
    if out_dir.is_file():
        sys.exit('Error: {} is a file (must be a directory or not exist)'.format(out_dir))
    if not out_dir.is_dir():
        try:
            out_dir.mkdir(parents=True)
            print('\nMade output directory: {}'.format(out_dir), flush=True)
        except OSError:
            sys.exit('Error: could not make directory {}'.format(out_dir))
