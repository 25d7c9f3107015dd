#!/usr/bin/env python
"""Copy the judged evidence of one gpurun capture (tools/prof.sh <tag>) into profiles/:
    python tools/make_profiles.py <tag> <round-label>        e.g.  r1o r1
writes profiles/launches_<label>.csv, ncu_pair_kernels_<label>.txt, ncu_source_<kernel>_<label>.txt,
k_pairs_traffic.json (read by bench.py for roofline.traffic) and bench_<label>.json if gpurun_out has one."""
import csv, io, json, os, shutil, subprocess, sys

tag, label = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
rep = os.path.join(G, f'prof_{tag}.ncu-rep')
shutil.copy(os.path.join(G, f'launches_{tag}.csv'), os.path.join(P, f'launches_{label}.csv'))
summary = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
open(os.path.join(P, f'ncu_pair_kernels_{label}.txt'), 'w').write(summary)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
detail, seen = {}, set()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d['Kernel Name'].split('(')[0].replace('void ', '')
    if name in seen:
        continue
    seen.add(name)
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    rd = float(d['dram__bytes_read.sum']) * scale[rows[1][hdr.index('dram__bytes_read.sum')]]
    wr = float(d['dram__bytes_write.sum']) * scale[rows[1][hdr.index('dram__bytes_write.sum')]]
    detail[name] = {'read': int(rd), 'write': int(wr), 'us': float(d['gpu__time_duration.sum'])}
pair = [k for k in detail if k.split('<')[0] in ('k_search', 'k_classify', 'k_hscan')]
json.dump({'atoms': 100000, 'dram_bytes_per_launch': sum(detail[k]['read'] + detail[k]['write'] for k in pair),
           'kernels': pair, 'detail': detail,
           'source': f'ncu --set full --clock-control none (caches flushed before every replayed kernel, so lists that stay in L2 '
                     f'between the kernels of a real step are read from DRAM here), capture {tag}, profiles/ncu_pair_kernels_{label}.txt'},
          open(os.path.join(P, 'k_pairs_traffic.json'), 'w'), indent=1)
for k in detail:
    base = k.split('<')[0]
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), rep, base, '45'], capture_output=True, text=True).stdout
    out2 = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), rep, base, '15', 'samples'], capture_output=True, text=True).stdout
    open(os.path.join(P, f'ncu_source_{base}_{label}.txt'), 'w').write(
        '# top source lines by warp instructions executed\n' + out + '\n# top source lines by stall samples\n' + out2)
b = os.path.join(G, f'bench_{tag}.json')
if os.path.exists(b):
    shutil.copy(b, os.path.join(P, f'bench_{label}.json'))
print('profiles/ updated from', tag, sorted(detail))
