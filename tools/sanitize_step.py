#!/usr/bin/env python
"""One small run of every kernel family for compute-sanitizer (run under gpurun):
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.engine import ContactEngine  # noqa: E402
from arpeggio_b200.soa import AtomSoA  # noqa: E402

p = params.make_params()
with ContactEngine(0, p) as eng:
    for n in (3000, 12000):
        soa = synth.cloud_featured(n, seed=5 + n)
        for _ in range(2):
            rec = eng.pairs(soa)
        eng.atom_sifts()
        rings, amides = synth.plane_set(64, 256, n_atoms=n, seed=9)
        eng.upload_planes(rings, amides)
        eng.ring_ring(); eng.atom_ring(); eng.amide_amide(); eng.amide_ring()
        print(n, 'atoms', rec.shape[0], 'records')
    batch = AtomSoA.concat([synth.cloud_featured(k, seed=40 + k) for k in (700, 0, 5, 1500)])
    print('batch', eng.pairs(batch).shape[0], 'records')
