#!/usr/bin/env python
"""The configs[2] cloud with every second atom an explicit hydrogen (a hydrogenated structure: ~3/4 of the within-cutoff pairs
involve a hydrogen and are dropped by interactions.py:712-713): step time and record count (A/B tooling, never a bench number).
    [ARPEGGIO_CUDA_LIB=variants/lib_x.so] python tools/hydrogen_ab.py <label>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arpeggio_b200 import abi, params, synth
from arpeggio_b200.engine import ContactEngine
soa = synth.cloud_featured(100000, seed=2)
soa.feat[1::2] |= abi.F_ELEM_H
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa); n = eng.run_pairs()
    eng.time_pairs(20, flush_l2=True)
    ms = eng.time_pairs(200, flush_l2=True); st = eng.stats()
    print('%-6s 50 %% hydrogens: us/step %6.1f | search %5.1f classify %5.1f hscan %5.1f | %d records' % (
        sys.argv[1] if len(sys.argv) > 1 else 'cur', ms * 1e3, st['ms_search'] * 1e3, (st['ms_classify'] - st['ms_hscan']) * 1e3, st['ms_hscan'] * 1e3, n))
