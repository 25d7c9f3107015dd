#!/usr/bin/env python
"""Phase times of the register grid kernel (diagnostic; needs a library built with -DGRID_PROFILE; run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
soa = synth.cloud_featured(atoms, seed=2)
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa)
    for _ in range(6):
        eng.run_pairs()
    eng.time_pairs(4, flush_l2=True)
