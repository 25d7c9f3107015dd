#!/usr/bin/env python
"""A few calls of arp_planes_run_all on the configs[3] inputs for ncu / wall timing (never a bench number)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
import ctypes as C
p = params.make_params()
soa = synth.cloud_featured(100000, seed=2)
rings, amides = synth.plane_set(2048, 12_500, n_atoms=100000, seed=3)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
with ContactEngine(0, p) as eng:
    eng.upload_atoms(soa); eng.upload_planes(rings, amides)
    n = (C.c_uint64 * 4)()
    for _ in range(3):
        eng._check(eng._L.arp_planes_run_all(eng._ctx, n))
    t0 = time.perf_counter()
    for _ in range(reps):
        eng._check(eng._L.arp_planes_run_all(eng._ctx, n))
    print('run_all alone: %.1f us per call' % ((time.perf_counter() - t0) / reps * 1e6), list(n))
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.planes_all()
    print('planes_all (run + 4 fetches into fresh arrays): %.1f us per call' % ((time.perf_counter() - t0) / reps * 1e6))
