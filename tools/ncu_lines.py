#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep:
    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep k_search [top] [samples]"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by_samples = len(sys.argv) > 4 and sys.argv[4] == 'samples'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kern],
                     capture_output=True, text=True).stdout
seen_launch = 0
rows, path, hdr = [], '', None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == 'File Path':
        path = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Kernel Name':
        seen_launch += 1
        if seen_launch > 1:
            break
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and r[0]:
        d = dict(zip(hdr[2:], r[2:]))
        try:
            rows.append((int(d['Instructions Executed']), int(d['# Samples'] or 0), float(d['Avg. Threads Executed'] or 0),
                         path, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(r[0] for r in rows) or 1
ts = sum(r[1] for r in rows) or 1
print(f'{kern}: {tot} warp instructions, {ts} samples, {len(rows)} source lines')
for n, s, act, p, ln, src in sorted(rows, key=lambda r: (r[1], r[0]) if by_samples else (r[0], r[1]), reverse=True)[:top]:
    print(f'{n:9d} {100 * n / tot:5.1f}% samp={100 * s / ts:5.1f}% act={act:4.1f} {p}:{ln} | {src[:100]}')
