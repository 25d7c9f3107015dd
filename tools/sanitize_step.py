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
        got = eng.planes_all()                                  # the grid path, all four terms in one sequence
        print(n, 'atoms', rec.shape[0], 'records;', {k: v.shape[0] for k, v in got.items()}, 'plane records')
        eng.run_pairs_async()                                   # compact sorted stream + distances on demand
        cp = eng.fetch_pairs_compact(with_dist=False)
        eng.fetch_pairs_dist(cp.n)
    batch = AtomSoA.concat([synth.cloud_featured(k, seed=40 + k) for k in (700, 0, 5, 1500)])
    print('batch', eng.pairs(batch).shape[0], 'records')
    parts = [synth.cloud_featured(k, seed=60 + k) for k in (900, 5, 2500, 1200)]
    off = eng.upload_atoms_batch(parts)                         # device-side concatenation
    eng.run_pairs_async()
    print('device-packed batch', eng.fetch_pairs_compact(with_dist=True).n, 'records', off.tolist())
    big = synth.cloud_featured(24000, seed=77)                  # binding-site flags through the cell grid (>= 20 000 atoms)
    eng.upload_atoms(big)
    print('flag_within (grid)', int(eng.flag_within(6.0).sum()), 'atoms flagged')
    # wire-form uploads (counts, sparse neighbours, fixed-point hydrogens), the packed view built behind the run
    from arpeggio_b200.engine import PackedPairs  # noqa: E402
    import numpy as np  # noqa: E402
    for n, dec in ((9000, 3), (5000, None), (140_000, 3)):
        src = synth.cloud_featured(n, seed=90 + n, h_decimals=dec)
        eng.upload_atoms(src.to_wire())
        eng.run_pairs_async()
        out = PackedPairs(np.zeros(n + 2, np.uint32), np.zeros(16 * n, np.uint32), np.zeros(16 * n, np.uint8), np.zeros(16 * n, np.float32))
        eng.fetch_pairs_packed_async(out, 5 * n, True)              # fewer words than the run has: the wait fetches the rest
        pk = eng.fetch_pairs_packed_wait()
        print('wire + packed', n, pk.n, 'records', pk.bits_j, 'bits,', pk.n_faults, 'fault records')
    off = eng.upload_atoms_batch([q.to_wire() if k % 2 else q for k, q in enumerate(parts)])
    eng.run_pairs_async()
    out = PackedPairs(np.zeros(int(off[-1]) + 2, np.uint32), np.zeros(100_000, np.uint32), None, None)
    eng.fetch_pairs_packed_async(out, 0, False)
    print('device-packed batch, mixed forms, packed view', eng.fetch_pairs_packed_wait().n, 'records')
if os.environ.get('ARPEGGIO_TILES'):
    print('the fused tile kernel (k_tiles) ran in place of k_search + k_classify')
