#!/usr/bin/env python
"""End-to-end time per 100k-atom structure when several of them share one launch sequence (BatchRunner.run(packed=True, pack=k)):
the merge kernels cost what the longer kernels save.  Run under gpurun."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.batch import BatchRunner
from arpeggio_b200.engine import pinned_soa
p = params.make_params()
soa = synth.cloud_featured(100_000, seed=2)
host = pinned_soa(soa.to_wire())
for slots in (4, 8):
    with BatchRunner(device=0, slots=slots, params=p) as runner:
        for pack in (1, 2, 3, 4, 8):
            n = 480
            runner.run([host] * (2 * slots * pack), check_finite=False, packed=True, pack=pack)
            best = min(runner.run([host] * n, check_finite=False, packed=True, pack=pack)[1] for _ in range(3)) / n
            print(f'slots={slots} pack={pack}: {best * 1e3:.4f} ms per structure', flush=True)
