#!/usr/bin/env python
"""Device step and its grid / search / classify / hscan split (events between the kernels) by structure size.  Run under gpurun."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
p = params.make_params()
with ContactEngine(0, p) as eng:
    for n in (100_000, 300_000, 1_000_000, 3_000_000):
        soa = synth.cloud_featured(n, seed=5, bonds=(n <= 1_000_000))
        eng.upload_atoms(soa)
        npairs = eng.run_pairs()
        eng.time_pairs(3, True)
        ms = eng.time_pairs(10, True)
        st = eng.stats()
        print(n, npairs, 'step %.1f us' % (ms * 1e3), {k: round(st[k] * 1e3, 1) for k in ('ms_grid', 'ms_search', 'ms_classify', 'ms_hscan', 'ms_pairs')},
              '%.2e pairs/s' % (npairs / ms * 1e3), flush=True)
