#!/usr/bin/env python
"""A few steps of the configs[2] job for ncu (never a bench number): python tools/profile_step.py [atoms] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.engine import ContactEngine  # noqa: E402

atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
soa = synth.cloud_featured(atoms, seed=2)
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa)
    for _ in range(steps):
        n = eng.run_pairs()
    print(atoms, 'atoms', n, 'records', eng.stats())
