#!/usr/bin/env python3
"""Map ncu's per-SASS-instruction stall samples back to CUDA source lines.
Usage: python tools/ncu_hotlines.py <report.ncu-rep> <library.so> <kernel-substring> [topN]"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, lib, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if 'Address' in r and 'Source' in r)
    start = rows.index(hdr) + 1
    ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    insts = []
    for r in rows[start:]:
        if len(r) <= isamp or not r[ia].startswith('0x'):
            break
        insts.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
    base = insts[0][0]
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    line_of = {}
    for f in os.listdir(tmp):
        if not f.endswith('.cubin'):
            continue
        txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        cur_fn, cur_line, active = None, None, False
        for ln in txt.splitlines():
            m = re.match(r'\s*\.text\.(\S+):', ln) or re.match(r'^(\S+):\s*$', ln)
            if ln.startswith('//--------------------- .text.'):
                active = kern in ln
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
            if m and cur_line:
                line_of[int(m.group(1), 16)] = cur_line
        if line_of:
            break
    agg = collections.Counter()
    exe = collections.Counter()
    total = 0
    for addr, src, samp, ex in insts:
        key = line_of.get(addr - base, ('?', 0))
        agg[key] += samp
        exe[key] += ex
        total += samp
    print('total samples', total, 'instructions', len(insts), 'mapped', len(line_of))
    srcs = {}
    for (f, l), s in agg.most_common(top):
        text = ''
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', f)
        if os.path.exists(path):
            if path not in srcs:
                srcs[path] = open(path).read().splitlines()
            if 0 < l <= len(srcs[path]):
                text = srcs[path][l - 1].strip()[:90]
        print('{:6d} {:5.1f}%  exec {:9d}  {}:{}  {}'.format(s, 100.0 * s / max(total, 1), exe[(f, l)], f, l, text))


if __name__ == '__main__':
    main()
