#!/usr/bin/env python3
"""Attribute the warp-state samples of one kernel in an ncu report to CUDA source lines.

    python tools/ncu_lines.py report.ncu-rep <kernel substring> <object>.cu [top]

ncu's --page source CSV lists SASS with samples but no line numbers; nvdisasm -g on
the cubin extracted from the in-tree library supplies the address -> line map."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'vittles_b200', 'lib', 'libvittles_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, check=True, capture_output=True)
cubin = os.path.join(tmp, src.replace('.cu', '') + '.sm_100a.cubin')
sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
off2line, cur, inside = {}, None, False
for l in sass:
    if l.startswith('.text.'):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr, data = rows[hi], [r for r in rows[hi + 1:] if r and r[0].startswith('0x')]
ia, isamp, iex = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
base = int(data[0][ia], 16)
agg, agx, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    ln = off2line.get(int(r[ia], 16) - base)
    agg[ln] += int(r[isamp]); agx[ln] += int(r[iex]); tot += int(r[isamp])
text = open(os.path.join(root, 'vittles_b200', 'csrc', src)).read().splitlines()
print('total samples', tot)
for ln, c in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    s = text[ln[1] - 1].strip()[:90] if ln and ln[0] == src else ''
    print('%-28s %6d %5.1f%% exec %8d  %s' % (ln, c, 100.0 * c / max(tot, 1), agx[ln], s))
