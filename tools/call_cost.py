#!/usr/bin/env python
"""Host-side cost of the three API calls of one structure (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arpeggio_b200 import abi, params, synth
from arpeggio_b200.engine import ContactEngine, PinnedBuffer, pinned_soa
atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
soa = pinned_soa(synth.cloud_featured(atoms, seed=1000))
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa); n = eng.run_pairs()
    pin = PinnedBuffer(16 * (n + 1024)); out = pin.array(abi.PAIR_DTYPE)
    t = np.zeros(4)
    R = 300
    for _ in range(R):
        t0 = time.perf_counter(); eng.upload_atoms(soa, check_finite=False)
        t1 = time.perf_counter(); eng.sync()
        t2 = time.perf_counter(); n = eng.run_pairs()
        t3 = time.perf_counter(); eng.fetch_pairs(n, sorted=False, out=out)
        t4 = time.perf_counter()
        t += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
    print(f'{atoms} atoms, {n} records: upload call {t[0]/R*1e6:.0f} us, upload drain {t[1]/R*1e6:.0f} us, run_pairs {t[2]/R*1e6:.0f} us, '
          f'fetch {t[3]/R*1e6:.0f} us; kernels {eng.stats()["ms_total"]*1e3:.0f} us')
