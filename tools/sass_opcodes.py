#!/usr/bin/env python3
"""Per-kernel count of the SASS opcodes that prove the Blackwell-native path (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk),
UTMALDG/UTMASTG (tensor-map TMA), SYNCS (mbarrier), HMMA (legacy mma.sync), LDGSTS (cp.async).
Usage: python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / 'deepbinner_b200' / 'libdeepbinner_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCCP', 'UBLKCP', 'UTMALDG', 'UTMASTG',
       'SYNCS', 'HMMA', 'HGMMA', 'LDGSTS', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'SHFL', 'FFMA', 'DFMA']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', cur)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            op = m.group(1)
            counts[cur]['total'] += 1
            for o in OPS:
                if op == o or op.startswith(o + '.') or (o in ('UTCHMMA',) and op.startswith(o)):
                    counts[cur][o] += 1
    print('SASS opcode counts per kernel of', pathlib.Path(LIB).name, '(cuobjdump -sass; sm_100a)')
    print('tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP,')
    print('cp.async.bulk.tensor -> UTMALDG/UTMASTG (not used: the weights are 1-D bulk copies, the inputs plain loads),')
    print('mbarrier -> SYNCS, mma.sync would be HMMA (none)')
    print()
    hdr = ['total'] + OPS
    print('{:60s} '.format('kernel') + ' '.join('{:>7s}'.format(h) for h in hdr))
    for k, c in counts.items():
        print('{:60s} '.format(k[:60]) + ' '.join('{:7d}'.format(c[h]) for h in hdr))


if __name__ == '__main__':
    main()
