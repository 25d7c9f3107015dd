#!/usr/bin/env python
"""Randomised parity stress (run under gpurun; not part of the test-suite): many seeded structures of random size, density,
cutoff, hydrogen fraction and bond density through every path -- default pair kernels, device-packed batches, compact stream,
plane grids, binding-site flags -- each compared bit for bit with the CPU oracle.
    python tools/stress.py [cases] [first seed]        (ARPEGGIO_TILES=1 for the fused tile kernel)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import util
from arpeggio_b200 import abi, params as arp_params, synth
from arpeggio_b200.engine import ContactEngine, PackedPairs
from oracle import oracle

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time()
checked = dict(pairs=0, records=0, batches=0, planes=0, within=0, wire_packed=0, h_fix=0)


def make(rng, n, seed):
    soa = synth.cloud_featured(n, seed=seed, atoms_per_residue=int(rng.integers(1, 14)), chain_len=int(rng.integers(2, 60)),
                               bonds=bool(rng.integers(2)))
    scale = float(rng.choice([0.5, 0.8, 1.0, 1.0, 1.6, 3.0]))
    shift = float(rng.choice([0.0, 0.0, -350.0, 7000.0]))
    soa.xyz[:] = np.round(soa.xyz.astype(np.float64) * scale + shift, 3).astype(np.float32)
    owner = np.repeat(np.arange(n), np.diff(soa.h_off))
    d = rng.normal(size=(owner.shape[0], 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    soa.h_xyz[:] = soa.xyz[owner].astype(np.float64) + d * rng.uniform(0.6, float(rng.choice([1.1, 2.2])), size=(owner.shape[0], 1))
    if soa.xnbr_xyz is not None:
        soa.xnbr_xyz[:] = (soa.xnbr_xyz.astype(np.float64) * scale + shift).astype(np.float32)
    if rng.random() < 0.5:                                  # explicit hydrogens: kept out of the cell grid
        soa.feat[rng.random(n) < float(rng.choice([0.1, 0.5]))] |= abi.F_ELEM_H
    return soa


with ContactEngine(0) as eng:
    for c in range(cases):
        rng = np.random.default_rng(50_000 + seed0 + c)
        p = arp_params.make_params(float(rng.choice([3.0, 4.0, 5.0, 5.0, 6.5, 8.0])), float(rng.choice([0.0, 0.1, 0.1, 0.4])), bool(rng.integers(2)))
        eng.set_params(p)
        n = int(rng.choice([1, 2, 17, 150, 900, 3000, 8000, 21000, 40000]))
        soa = make(rng, n, 60_000 + seed0 + c)
        exp = oracle.pairs(soa, p)
        util.assert_records_equal(eng.pairs(soa), exp, f'case {c}: pairs n={n}')
        eng.run_pairs_async()
        cp = eng.fetch_pairs_compact(with_dist=True)
        util.assert_records_equal(cp.to_records(), exp, f'case {c}: compact')
        checked['pairs'] += 1; checked['records'] += exp.shape[0]
        # wire-form upload, packed view built behind the run, a blind copy of a random share of the words
        if rng.random() < 0.5:
            soa.h_xyz[:] = np.round(soa.h_xyz, 3)
            exp = oracle.pairs(soa, p)
        w = soa.to_wire()
        checked['h_fix'] += w.h_fix is not None
        eng.upload_atoms(w)
        eng.run_pairs_async()
        cap = exp.shape[0] + int(rng.integers(0, 50))
        out = PackedPairs(np.zeros(n + 2, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.float32))
        eng.fetch_pairs_packed_async(out, int(rng.choice([0, 1, exp.shape[0] // 2, exp.shape[0], 10 ** 9])), True)
        pk = eng.fetch_pairs_packed_wait()
        util.assert_records_equal(pk.to_records(soa.feat), exp, f'case {c}: wire + packed')
        checked['wire_packed'] += 1
        r = float(rng.choice([6.0, 2.0, 11.0]))
        assert np.array_equal(eng.flag_within(r), oracle.flag_within(soa, r)), f'case {c}: within {r}'
        checked['within'] += 1
        if n >= 150 and c % 2 == 0:
            rings, amides = synth.plane_set(max(2, n // 40), max(2, n // 8), n_atoms=n, seed=70_000 + c)
            scale = float(np.abs(soa.xyz).max()) / max(float(np.abs(rings.center).max()), 1.0)
            eng.upload_atoms(soa); eng.upload_planes(rings, amides)
            got = eng.planes_all()
            util.assert_records_equal(got['ring_ring'], oracle.ring_ring(rings, p), f'case {c}: ring-ring')
            util.assert_records_equal(got['atom_ring'], oracle.atom_ring(soa, rings, p), f'case {c}: atom-ring')
            util.assert_records_equal(got['amide_amide'], oracle.amide_amide(amides, p), f'case {c}: amide-amide')
            util.assert_records_equal(got['amide_ring'], oracle.amide_ring(amides, rings, p), f'case {c}: amide-ring')
            checked['planes'] += 1
        if c % 3 == 0:
            parts = [make(rng, int(rng.choice([0, 1, 40, 700, 2500, 6000])), 80_000 + 10 * c + k) for k in range(int(rng.integers(2, 7)))]
            parts = [q for q in parts if q.n_atoms > 0] or [make(rng, 30, 81_000 + c)]
            off = eng.upload_atoms_batch([q.to_wire() if (k + c) % 2 else q for k, q in enumerate(parts)])
            eng.run_pairs_async()
            cb = eng.fetch_pairs_compact(with_dist=True)
            pb = eng.fetch_pairs_packed(with_dist=True)
            for k, part in enumerate(parts):
                rec = cb.structure(int(off[k]), int(off[k + 1])).to_records()
                rec['j'] -= int(off[k])
                want = oracle.pairs(part, p)
                util.assert_records_equal(rec, want, f'case {c}: packed structure {k}')
                util.assert_records_equal(pb.structure(int(off[k]), int(off[k + 1])).to_records(part.feat), want, f'case {c}: packed words of structure {k}')
            checked['batches'] += 1
print('stress ok:', checked, 'in %.0f s' % (time.time() - t0), '(tile kernel)' if os.environ.get('ARPEGGIO_TILES') else '')
