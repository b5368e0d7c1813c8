"""Summary table of barcode calls - same output format as reference `misc.py:19-36`."""

import collections
import sys


def print_summary_table(classifications, output=sys.stderr):
    tally = collections.Counter(classifications.values())
    numeric = sorted((int(b), b) for b in tally if b.isdigit())
    other = sorted(b for b in tally if not b.isdigit())
    print('', file=output)
    print('Barcode     Count', file=output)
    for barcode in [b for _, b in numeric] + other:
        print('{:>7} {:>9}'.format(barcode, tally[barcode]), file=output)
    print('', file=output)
