#!/usr/bin/env python
"""configs[3] timing: the four plane terms on the synthetic plane set (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
A = int(sys.argv[3]) if len(sys.argv) > 3 else 12500
soa = synth.cloud_featured(atoms, seed=2)
rings, amides = synth.plane_set(R, A, n_atoms=atoms, seed=3)
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa); eng.upload_planes(rings, amides)
    for name in ('ring_ring', 'atom_ring', 'amide_amide', 'amide_ring'):
        f = getattr(eng, name)
        n = f().shape[0]
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        dt = (time.perf_counter() - t0) / 10
        print(f'{name:12s} {n:8d} records  {dt * 1e6:9.0f} us per call (run + fetch)')
