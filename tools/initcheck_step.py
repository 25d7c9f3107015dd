#!/usr/bin/env python
"""A small run of the pair path, the wire decode and the sorted views for compute-sanitizer --tool initcheck."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.engine import ContactEngine, PackedPairs  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
with ContactEngine(0, params.make_params()) as eng:
    src = synth.cloud_featured(3000, seed=7, h_decimals=3)
    if which in ('all', 'plain'):
        rec = eng.pairs(src)
        print('plain', rec.shape[0])
    if which in ('all', 'wire'):
        eng.upload_atoms(src.to_wire())
        eng.run_pairs_async()
        out = PackedPairs(np.zeros(3002, np.uint32), np.zeros(60000, np.uint32), None, np.zeros(60000, np.float32))
        eng.fetch_pairs_packed_async(out, 0, True)
        print('wire + packed', eng.fetch_pairs_packed_wait().n)
    if which in ('all', 'batch'):
        parts = [synth.cloud_featured(k, seed=60 + k) for k in (900, 5, 2500)]
        eng.upload_atoms_batch([q.to_wire() if k % 2 else q for k, q in enumerate(parts)])
        eng.run_pairs_async()
        print('batch', eng.fetch_pairs_packed(with_dist=False).n)
