#!/usr/bin/env python
"""A few calls of arp_planes_run_all on the configs[3] inputs for ncu / wall timing (never a bench number)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
import ctypes as C
p = params.make_params()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
atoms = int(sys.argv[2]) if len(sys.argv) > 2 else 100000          # rings and amides scale with the atoms (configs[3] ratios)
soa = synth.cloud_featured(atoms, seed=2, bonds=False)
rings, amides = synth.plane_set(2048 * atoms // 100000, 12_500 * atoms // 100000, n_atoms=atoms, seed=3)
with ContactEngine(0, p) as eng:
    eng.upload_atoms(soa); eng.upload_planes(rings, amides)
    n = (C.c_uint64 * 4)()
    for _ in range(3):
        eng._check(eng._L.arp_planes_run_all(eng._ctx, n))
    t0 = time.perf_counter()
    for _ in range(reps):
        eng._check(eng._L.arp_planes_run_all(eng._ctx, n))
    print('%d atoms, %d rings, %d amides | run_all alone: %.1f us per call' % (atoms, rings.n, amides.n, (time.perf_counter() - t0) / reps * 1e6), list(n))
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.planes_all()
    print('planes_all (run + 4 fetches into fresh arrays): %.1f us per call' % ((time.perf_counter() - t0) / reps * 1e6))
