"""
fast5 discovery and raw-signal extraction with the semantics of reference `load_fast5s.py`
(`get_read_id_and_signal` :25-49, `find_all_fast5s` :52-64, `determine_single_or_multi_fast5s`
:67-90, `get_root_level_keys` :93-98), on top of the built-in HDF5 subset reader (no h5py here).
"""

import os
import random
import sys

from . import hdf5_lite


def get_read_id_and_signal(fast5_file):
    """-> (read_id str, int16 signal) or (None, None) if the file cannot be read."""
    try:
        with hdf5_lite.open_file(fast5_file) as h:
            keys = h.keys()
            if 'Raw' in keys:           # old single-read layout: /Raw/Reads/Read_<n>
                group = h['Raw/Reads'].values()[0]
            else:                       # new layout: /read_<uuid>/Raw
                reads = [k for k in keys if k.startswith('read_')]
                if len(reads) > 1:
                    sys.exit('Error: Deepbinner does not (yet) support multi-read fast5 files')
                if not reads:
                    return None, None
                group = h[reads[0] + '/Raw']
            read_id = group.attrs['read_id']
            if isinstance(read_id, bytes):
                read_id = read_id.decode()
            signal = group['Signal'].read()
        return str(read_id), signal
    except (OSError, KeyError, IndexError, ValueError):
        return None, None


def find_all_fast5s(directory, verbose=False):
    if verbose:
        print('Looking for fast5 files in {}... '.format(directory), file=sys.stderr, end='',
              flush=True)
    found = [os.path.join(root, name)
             for root, _, names in os.walk(str(directory)) for name in names
             if name.endswith('.fast5')]
    if verbose:
        print('{} {} found'.format(len(found), 'fast5' if len(found) == 1 else 'fast5s'),
              file=sys.stderr)
    return found


def get_root_level_keys(fast5_file):
    try:
        with hdf5_lite.open_file(fast5_file) as h:
            return h.keys()
    except (OSError, KeyError, IndexError, ValueError):
        return []


def determine_single_or_multi_fast5s(fast5s):
    """Inspect up to five randomly chosen files -> 'single' or 'multi' (exits on an old/multi mix)."""
    sample = list(fast5s)
    random.shuffle(sample)
    kinds = set()
    for path in sample[:5]:
        keys = get_root_level_keys(path)
        if 'Raw' in keys:
            kinds.add('single-old')
            continue
        n_reads = sum(1 for k in keys if k.startswith('read_'))
        if n_reads == 1:
            kinds.add('single-new')
        elif n_reads > 1:
            kinds.add('multi')
    if 'multi' in kinds and 'single-old' in kinds:
        sys.exit('Error: your reads appear to be a mixture of old and new formats. Deepbinner '
                 'can handle one or the other, but not both at once.')
    return 'multi' if 'multi' in kinds else 'single'
