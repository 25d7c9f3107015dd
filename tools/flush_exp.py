#!/usr/bin/env python
"""Experiment: per-kernel split with and without the L2 flush between steps (understanding, not a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
soa = synth.cloud_featured(atoms, seed=2)
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa)
    n = eng.run_pairs()
    for flush in (True, False):
        eng.time_pairs(20, flush)
        ms = eng.time_pairs(200, flush)
        st = eng.stats()
        print(f'atoms={atoms} pairs={n} flush={flush}: total {ms*1e3:.1f} us  grid {st["ms_grid"]*1e3:.1f}  search {st["ms_search"]*1e3:.1f}  classify {st["ms_classify"]*1e3:.1f}')
