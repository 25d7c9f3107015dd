"""GPU parity of the ring / amide plane terms against the golden records of the reference's own
code and against the CPU oracle on the synthetic plane set (BASELINE configs[3])."""
import numpy as np
import pytest

import util
from arpeggio_b200 import abi, params as arp_params, synth
from oracle import oracle

pytestmark = pytest.mark.gpu

CASES = [c for c in util.golden_cases() if c != 'xbond_fault']


@pytest.mark.parametrize('case', CASES)
def test_golden_planes(engine, case):
    g = util.Golden(case)
    engine.set_params(g.params)
    engine.upload_atoms(g.soa)
    engine.upload_planes(g.rings, g.amides)
    util.assert_records_equal(engine.ring_ring(), g.exp_ring_ring, f'{case} ring-ring')
    util.assert_records_equal(engine.atom_ring(), g.exp_atom_ring, f'{case} atom-ring')
    util.assert_records_equal(engine.amide_amide(), g.exp_amide_amide, f'{case} amide-amide')
    util.assert_records_equal(engine.amide_ring(), g.exp_amide_ring, f'{case} amide-ring')


def test_config4_synthetic_plane_set(engine):
    p = arp_params.make_params()
    engine.set_params(p)
    n_atoms = 100_000
    soa = synth.cloud_featured(n_atoms, seed=2, bonds=False)
    rings, amides = synth.plane_set(2048, 12_500, n_atoms, seed=3)
    engine.upload_atoms(soa)
    engine.upload_planes(rings, amides)
    rr = engine.ring_ring()
    util.assert_records_equal(rr, oracle.ring_ring(rings, p), 'ring-ring')
    assert rr.shape[0] > 100
    ar = engine.atom_ring()
    util.assert_records_equal(ar, oracle.atom_ring(soa, rings, p), 'atom-ring')
    assert ar.shape[0] > 100
    aa = engine.amide_amide()
    util.assert_records_equal(aa, oracle.amide_amide(amides, p), 'amide-amide')
    assert aa.shape[0] > 100
    util.assert_records_equal(engine.amide_ring(), oracle.amide_ring(amides, rings, p), 'amide-ring')


def test_dense_planes_hit_every_geometry(engine):
    """Rings packed into a small box: every one of the 9 geometries and two-label records occur."""
    p = arp_params.make_params()
    engine.set_params(p)
    rings, amides = synth.plane_set(400, 600, n_atoms=400, seed=5, n_residues=50)
    engine.upload_planes(rings, amides)
    rr = engine.ring_ring()
    util.assert_records_equal(rr, oracle.ring_ring(rings, p), 'ring-ring dense')
    first = set((rr['code'] & 0xF).tolist())
    assert first >= set(range(9))
    assert np.any(((rr['code'] >> 4) & 0xF) != 0xF)
    util.assert_records_equal(engine.amide_amide(), oracle.amide_amide(amides, p), 'amide-amide dense')
    util.assert_records_equal(engine.amide_ring(), oracle.amide_ring(amides, rings, p), 'amide-ring dense')


def test_degenerate_planes(engine):
    """Zero normals (NaN cosines), coincident centres, unit-cosine planes."""
    p = arp_params.make_params()
    engine.set_params(p)
    rings, amides = synth.plane_set(64, 64, n_atoms=100, seed=6, n_residues=8)
    rings.normal[:8] = 0.0
    rings.center[8:16] = rings.center[8]
    rings.normal[16:24] = rings.normal[16]
    amides.normal[:8] = 0.0
    amides.center[8:16] = amides.center[8]
    amides.normal[16:24] = amides.normal[16]
    amides.center[16:24] = amides.center[16] + amides.normal[16] * np.arange(8, dtype=np.float32)[:, None] * np.float32(0.5)
    engine.upload_planes(rings, amides)
    util.assert_records_equal(engine.ring_ring(), oracle.ring_ring(rings, p), 'ring-ring degenerate')
    util.assert_records_equal(engine.amide_amide(), oracle.amide_amide(amides, p), 'amide-amide degenerate')
    util.assert_records_equal(engine.amide_ring(), oracle.amide_ring(amides, rings, p), 'amide-ring degenerate')


def test_empty_planes(engine):
    from arpeggio_b200.soa import PlaneSoA
    engine.upload_planes(PlaneSoA.empty(False), PlaneSoA.empty(True))
    assert engine.ring_ring().shape[0] == 0
    assert engine.amide_amide().shape[0] == 0
    assert engine.amide_ring().shape[0] == 0


@pytest.mark.parametrize('case', CASES)
def test_ring_assignment_golden(engine, case):
    """ring -> residue assignment as the reference's own function left it."""
    g = util.Golden(case)
    engine.set_params(g.params)
    f = g.f4
    atom, dist = engine.ring_nearest_atom(f['xyz'], f['centers'], 3.0)
    res = np.where(atom >= 0, f['atom_res'][np.maximum(atom, 0)], -1)
    assert np.array_equal(res, f['ring_res'])
    assert np.array_equal(dist.view(np.uint64), f['ring_dist'].view(np.uint64))
    o_atom, o_dist = oracle.ring_nearest_atom(f['xyz'], f['centers'], 3.0, g.params)
    assert np.array_equal(atom, o_atom) and np.array_equal(dist.view(np.uint64), o_dist.view(np.uint64))


def test_ring_assignment_large_and_edges(engine):
    p = arp_params.make_params()
    engine.set_params(p)
    soa = synth.cloud_featured(100_000, seed=2)
    rings, _ = synth.plane_set(2048, 8, n_atoms=100_000, seed=3)
    for radius in (3.0, 0.9, 0.0):
        atom, dist = engine.ring_nearest_atom(soa.xyz, rings.center, radius)
        o_atom, o_dist = oracle.ring_nearest_atom(soa.xyz, rings.center, radius, p)
        assert np.array_equal(atom, o_atom) and np.array_equal(dist.view(np.uint64), o_dist.view(np.uint64))
    assert (atom == -1).any()
    # exact ties: two atoms mirror-symmetric about the centroid -> lowest index; duplicates of one atom
    xyz = np.array([[1, 0, 0], [-1, 0, 0], [0, 2, 0], [0, 2, 0], [50, 50, 50]], np.float32)
    centers = np.array([[0, 0, 0], [0, 2.5, 0], [50, 50, 53.0000001], [50, 50, 53]], np.float64)
    atom, dist = engine.ring_nearest_atom(xyz, centers, 3.0)
    assert atom.tolist() == [0, 2, -1, 4] and dist.tolist() == [1.0, 0.5, 0.0, 3.0]
    # no atoms / no rings
    a, d = engine.ring_nearest_atom(np.zeros((0, 3), np.float32), centers, 3.0)
    assert a.tolist() == [-1] * 4
    a, d = engine.ring_nearest_atom(xyz, np.zeros((0, 3)), 3.0)
    assert a.shape == (0,)
