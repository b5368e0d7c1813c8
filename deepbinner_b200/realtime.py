"""
`deepbinner realtime`: watch a directory the sequencer writes fast5 files into, classify new files
on the GPU and move them into per-barcode directories.  Behaviour follows reference
`deepbinner/realtime.py` (`realtime` :28-70, `look_for_new_fast5s` :73-78, `classify_and_move`
:81-108, `move_classified_fast5s` :111-143, `get_directory_name` :146-150): 5 s polling, at most
20 000 single-read files per round, files whose destination already exists are ignored from then
on, `--stop` ends the loop when nothing is waiting, Ctrl-C exits cleanly.

Multi-read fast5 files: the reference shells out to ONT's `multi_to_single_fast5`; here they are
rejected with the reference's `classify` message (unpacking is outside the accelerated path).
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
POLL_SECONDS = 5


def realtime(args, poll_seconds=POLL_SECONDS):
    args.verbose = False
    set_tensorflow_threads(args)
    start_model, start_input_size, end_model, end_input_size, output_size, _ = \
        load_and_check_models(args.start_model, args.end_model, args.scan_size,
                              out_dest=sys.stdout, device=getattr(args, 'device', 0))
    in_dir = pathlib.Path(args.in_dir)
    out_dir = pathlib.Path(args.out_dir)
    make_output_dir(out_dir)
    nested_out_dir = is_inside(out_dir, in_dir)

    print('\nLooking for new fast5 files in {}'.format(in_dir), flush=True)
    ignore_files = set()
    waiting_dots = 0
    try:
        while True:
            fast5s = [f for f in look_for_new_fast5s(in_dir, out_dir, nested_out_dir)
                      if f not in ignore_files]
            if fast5s:
                if waiting_dots:
                    print('', flush=True)
                    waiting_dots = 0
                if determine_single_or_multi_fast5s(fast5s) == 'multi':
                    sys.exit('Error: deepbinner realtime on the B200 engine requires one-read-per-'
                             'file fast5s - convert with multi_to_single_fast5 before running')
                print('\nFound {:,} fast5 files'.format(len(fast5s)), flush=True)
                time.sleep(poll_seconds)   # let files that are still being written settle
                classify_and_move(fast5s, args, start_model, start_input_size, end_model,
                                  end_input_size, output_size, out_dir, ignore_files)
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
                      output_size, out_dir, ignore_files):
    if len(fast5s) > MAX_FILES_PER_ROUND:
        fast5s = fast5s[:MAX_FILES_PER_ROUND]
        print('Limiting this round to {:,} files'.format(MAX_FILES_PER_ROUND), flush=True)
    classifications, read_id_to_fast5_file = \
        classify_fast5_files(fast5s, start_model, start_input_size, end_model, end_input_size,
                             output_size, args, full_output=False, verified_single_read=True)
    print('', flush=True)
    move_classified_fast5s(classifications, read_id_to_fast5_file, out_dir, ignore_files)
    print_summary_table(classifications, output=sys.stdout)


def move_classified_fast5s(classifications, read_id_to_fast5_file, out_dir, ignore_files):
    moved = 0
    total = len(classifications)
    for read_id, barcode_call in classifications.items():
        source = read_id_to_fast5_file[read_id]
        dest_dir = pathlib.Path(out_dir) / get_directory_name(barcode_call)
        dest_dir.mkdir(parents=True, exist_ok=True)
        dest = dest_dir / pathlib.Path(source).name
        if dest.exists():
            ignore_files.add(source)
            continue
        shutil.move(source, str(dest))
        moved += 1
        print('\rMoving fast5s: {:,} / {:,}'.format(moved, total), end='', flush=True)
    print('', flush=True)
    if total and not moved:
        sys.exit('Error: no files could be moved (do they already exist in the output directory?)')


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
