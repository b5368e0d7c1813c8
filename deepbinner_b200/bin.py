"""
`deepbinner bin`: split a FASTA/FASTQ read file into one file per barcode using a classification
table made by `deepbinner classify` - the step that follows the accelerated path in the user
workflow `classify -> bin` (reference `deepbinner/bin.py`: `bin_reads` :26-32,
`load_classifications` :35-72, `make_output_dir` :75-89, `write_read_files` :109-162,
`get_sequence_file_type` :171-187, `print_summary_and_zip` :190-213).  Pure host-side text streaming.

Same inputs, console messages, output file names (`barcodeNN.fastq.gz`, `unclassified.fastq.gz`) and
`sys.exit('Error: ...')` behaviour as the reference.  One deliberate difference: a read whose ID is
missing from the classification table makes the reference crash with a KeyError (it has no output
file for its 'not found' class, bin.py:145-153); here such reads are counted, reported in the summary
and not written.
"""

import collections
import gzip
import os
import pathlib
import re
import shutil
import subprocess
import sys

UUID = re.compile(r'[0-9a-fA-F]{8}-[0-9a-fA-F]{4}-[0-9a-fA-F]{4}-[0-9a-fA-F]{4}-[0-9a-fA-F]{12}')
NOT_FOUND = 'not found'


def bin_reads(args):
    classifications = load_classifications(args.classes)
    class_names = sorted({class_to_class_names(c) for c in classifications.values()})
    input_type = get_sequence_file_type(args.reads)
    out_filenames = get_output_filenames(class_names, args.out_dir, input_type)
    make_output_dir(args.out_dir, out_filenames)
    write_read_files(args.reads, classifications, out_filenames, input_type)


def load_classifications(class_filename):
    """read_id -> int barcode or None ('none'); header and short lines are skipped."""
    print('\nLoading classifications...', end='', flush=True)
    if not pathlib.Path(class_filename).is_file():
        sys.exit('Error: {} does not exist'.format(class_filename))
    classifications = {}
    odd_ids = False
    with open(class_filename, 'rt') as table:
        for line in table:
            fields = line.strip().split('\t')
            if len(fields) < 2 or fields[0].lower() == 'read_id':
                continue
            read_id, call = fields[0], fields[1]
            if len(read_id) != 36 or not all(read_id[i] == '-' for i in (8, 13, 18, 23)):
                odd_ids = True
            if call == 'none':
                classifications[read_id] = None
                continue
            try:
                classifications[read_id] = int(call)
            except ValueError:
                sys.exit('Error: read {} has a non-integer bin of {}'.format(read_id, call))
    print(' done')
    if odd_ids:
        print('Warning: one or more read IDs did not conform to the expected format (UUID)')
    print('{:,} total classifications found'.format(len(classifications)))
    print()
    return classifications


def class_to_class_names(classification):
    if classification is None:
        return 'unclassified'
    if isinstance(classification, str):
        return classification
    return 'barcode{:02d}'.format(classification)


def get_output_filenames(class_names, out_dir, input_type):
    return collections.OrderedDict(
        (name, str(pathlib.Path(out_dir) / (name + '.' + input_type))) for name in class_names)


def make_output_dir(out_dir, out_filenames):
    out_dir = pathlib.Path(out_dir)
    if out_dir.is_file():
        sys.exit('Error: {} is an existing file'.format(out_dir))
    if not out_dir.is_dir():
        try:
            os.makedirs(str(out_dir), exist_ok=True)
            print('Making output directory: {}/'.format(out_dir))
        except OSError:
            sys.exit('Error: unable to create output directory {}'.format(out_dir))
    for filename in out_filenames.values():
        for candidate in (filename, filename + '.gz'):
            if pathlib.Path(candidate).exists():
                sys.exit('Error: {} already exists'.format(candidate))
    print()


def get_compression_type(filename):
    """'plain' or 'gz' from the magic bytes; bzip2 / zip inputs are rejected like the reference
    (misc.py:39-58)."""
    with open(filename, 'rb') as f:
        start = f.read(4)
    if start.startswith(b'\x1f\x8b\x08'):
        return 'gz'
    if start.startswith(b'\x42\x5a\x68'):
        sys.exit('Error: cannot use bzip2 format - use gzip instead')
    if start.startswith(b'\x50\x4b\x03\x04'):
        sys.exit('Error: cannot use zip format - use gzip instead')
    return 'plain'


def get_open_function(filename):
    return gzip.open if get_compression_type(filename) == 'gz' else open


def get_sequence_file_type(filename):
    if not pathlib.Path(filename).is_file():
        sys.exit('Error: could not find ' + filename)
    with get_open_function(filename)(filename, 'rt') as f:
        try:
            first = f.read(1)
        except UnicodeDecodeError:
            first = ''
    if first == '>':
        return 'fasta'
    if first == '@':
        return 'fastq'
    raise ValueError('Error: could not determine file format (should be fasta or fastq)')


def iter_records(handle, input_type):
    """Yield (header_line, whole_record_text); 2 lines per FASTA record, 4 per FASTQ record, as the
    reference assumes (bin.py:129-136)."""
    lines_per_record = 4 if input_type == 'fastq' else 2
    while True:
        header = handle.readline()
        if not header:
            return
        rest = [handle.readline() for _ in range(lines_per_record - 1)]
        yield header, header + ''.join(rest)


def write_read_files(reads_filename, classifications, out_filenames, input_type):
    bin_counts = collections.Counter()
    writers = {name: open(filename, 'wt') for name, filename in out_filenames.items()}
    count = 0
    try:
        with get_open_function(reads_filename)(reads_filename, 'rt') as reads:
            for header, record in iter_records(reads, input_type):
                if count % 100 == 0:
                    print_progress(count)
                count += 1
                found = UUID.search(header)
                if found is None:
                    sys.exit('Error: could not find read ID in header: {}'.format(header))
                read_id = found.group(0)
                class_name = class_to_class_names(classifications[read_id]) \
                    if read_id in classifications else NOT_FOUND
                bin_counts[class_name] += 1
                if class_name in writers:
                    writers[class_name].write(record)
    finally:
        for w in writers.values():
            w.close()
    print_progress(count, carriage_return=False)
    print('\n')
    print_summary_and_zip(bin_counts, out_filenames)


def print_progress(count, carriage_return=True):
    print('Writing reads: {:,} '.format(count), end='\r' if carriage_return else '')


def print_summary_and_zip(bin_counts, out_filenames):
    gz = 'pigz' if shutil.which('pigz') else 'gzip'
    print('Gzipping reads (with pigz):' if gz == 'pigz' else 'Gzipping reads:')
    print('  Barcode       Reads     File')
    class_names = list(out_filenames)
    if NOT_FOUND in bin_counts:
        class_names.append(NOT_FOUND)
    for class_name in class_names:
        gzipped = ''
        if class_name in out_filenames:
            subprocess.check_output([gz, out_filenames[class_name]])
            gzipped = out_filenames[class_name] + '.gz'
        display = 'none' if class_name == 'unclassified' else class_name
        print('  {:<9} {:>9}     {}'.format(display, bin_counts[class_name], gzipped))
    print()
