#!/usr/bin/env python
"""configs[4] experiment: structures/s of BatchRunner against the number of stream slots (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.batch import BatchRunner
atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
from arpeggio_b200.engine import pinned_soa
distinct = [synth.cloud_featured(atoms, seed=1000 + k) for k in range(8)]
if os.environ.get('PINNED', '1') == '1':
    distinct = [pinned_soa(s) for s in distinct]
shard = [distinct[k % 8] for k in range(n)]
for slots in (1, 2, 3, 4, 6, 8, 12):
    r = BatchRunner(0, slots, params.make_params())
    r.run(shard[:2 * slots], check_finite=False)
    best = 1e9
    for _ in range(3):
        counts, dt = r.run(shard, check_finite=False)
        best = min(best, dt)
    r.close()
    print(f'slots {slots:2d}: {n / best:8.0f} structures/s  {sum(counts) / best / 1e9:.2f} Gpairs/s  {best / n * 1e6:.0f} us/structure')
