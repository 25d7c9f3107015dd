#!/usr/bin/env python
"""SASS opcode histogram of the step's kernels from the built library (no GPU needed):
    python tools/sass_hist.py > profiles/sass_histogram_<round>.txt
Shows what the kernels are made of: UBLKCP / SYNCS (cp.async.bulk, mbarrier), LDG / LDS / STS / ATOMG / RED, FP64 ..."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'arpeggio_b200', 'libarpeggio_cuda.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
want = ('k_grid_reg', 'k_search', 'k_classify', 'k_hscan', 'k_tiles', 'k_sort_', 'k_pl_', 'k_merge_', 'k_wg_', 'k_sift', 'k_grid_fused', 'k_cnt_', 'k_wire_')
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0].replace('void ', '')
        cur = name if any(w in name for w in want) else None
        if cur:
            hist[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)', line)
    if cur and m:
        op = m.group(1)
        if op in ('LDG', 'STG', 'LD', 'ST', 'LDS', 'STS', 'ATOMG', 'ATOMS', 'RED', 'UBLKCP', 'SYNCS', 'LDGSTS', 'UTMALDG'):
            mods = [x for x in m.group(2).split('.') if x in ('128', '64', 'U16', 'U8', 'E', 'ADD', 'OR', 'MAX', 'S', 'G', 'ARRIVE', 'PHASECHK', 'EXCH')]
            op += ''.join('.' + x for x in mods if x != 'E')
        hist[cur][op] += 1
groups = (('bulk async copy / mbarrier', ('UBLKCP', 'SYNCS', 'UTMALDG', 'LDGSTS')), ('global load', ('LDG', 'LD.')), ('global store', ('STG', 'ST.')),
          ('shared', ('LDS', 'STS')), ('atomics', ('ATOMG', 'ATOMS', 'RED')), ('fp64', ('DADD', 'DMUL', 'DFMA', 'DSETP', 'F2F', 'MUFU')),
          ('fp32', ('FADD', 'FMUL', 'FFMA', 'FSETP', 'FMNMX')), ('warp', ('SHFL', 'VOTE', 'MATCH', 'REDUX', 'WARPSYNC')), ('grid sync / PDL', ('ACQBULK', 'PREEXIT', 'ERRBAR', 'MEMBAR', 'CCTL')))
for k, h in hist.items():
    tot = sum(h.values())
    print(f'== {k}: {tot} SASS instructions')
    for title, keys in groups:
        sel = [(op, n) for op, n in h.items() if any(op == q or op.startswith(q) for q in keys)]
        if sel:
            print(f'   {title:28s} ' + ', '.join(f'{op} {n}' for op, n in sorted(sel)))
    top = ', '.join(f'{op} {n}' for op, n in h.most_common(8))
    print(f'   most frequent               {top}')
