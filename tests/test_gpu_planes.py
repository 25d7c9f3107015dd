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


def _all_terms(eng, soa, rings, amides, p, what):
    eng.set_params(p)
    eng.upload_atoms(soa)
    eng.upload_planes(rings, amides)
    util.assert_records_equal(eng.ring_ring(), oracle.ring_ring(rings, p), what + ' ring-ring')
    util.assert_records_equal(eng.atom_ring(), oracle.atom_ring(soa, rings, p), what + ' atom-ring')
    util.assert_records_equal(eng.amide_amide(), oracle.amide_amide(amides, p), what + ' amide-amide')
    util.assert_records_equal(eng.amide_ring(), oracle.amide_ring(amides, rings, p), what + ' amide-ring')


@pytest.mark.parametrize('knob', [None, 'ARPEGGIO_NO_PLANE_SCREEN'])
def test_screened_and_plain_plane_loops(monkeypatch, knob):
    """The float32 distance screen + hit bitmask in front of the plane predicates must not change a record: far from
    the origin (large float32 rounding), thresholds hit exactly, non-finite centres, column counts that are not a
    multiple of the tile, with and without the screen."""
    from arpeggio_b200.engine import ContactEngine
    if knob:
        monkeypatch.setenv(knob, '1')
    p = arp_params.make_params()
    with ContactEngine(0, p) as eng:
        soa = synth.cloud_featured(3_000, seed=31, bonds=False)
        rings, amides = synth.plane_set(300, 1_500, n_atoms=3_000, seed=32, n_residues=375)
        _all_terms(eng, soa, rings, amides, p, 'plain')
        # 40 km from the origin: one float32 ulp is 4e-3 A there
        off = np.array([40000.0, -35000.0, 20000.0])
        far = synth.cloud_featured(3_000, seed=31, bonds=False)
        far.xyz[:] = (far.xyz.astype(np.float64) + off).astype(np.float32)
        far.h_xyz[:] = far.h_xyz + off
        rings_f, amides_f = synth.plane_set(300, 1_500, n_atoms=3_000, seed=32, n_residues=375)
        rings_f.center[:] = rings_f.center + off
        amides_f.center[:] = (amides_f.center.astype(np.float64) + off).astype(np.float32)
        _all_terms(eng, far, rings_f, amides_f, p, 'far')
        # centroid distances exactly on / one ulp around the 6 A thresholds, along x
        rings_k, amides_k = synth.plane_set(64, 64, n_atoms=100, seed=33, n_residues=8)
        for k in range(0, 60, 2):
            d = np.nextafter(6.0, [0.0, 6.0, 12.0][k // 2 % 3]) if k // 2 % 3 != 1 else 6.0
            rings_k.center[k] = [100.0 * k, 0.0, 0.0]
            rings_k.center[k + 1] = [100.0 * k + d, 0.0, 0.0]
            rings_k.normal[k] = rings_k.normal[k + 1] = [0.0, 0.0, 1.0]
            d32 = [np.nextafter(np.float32(6.0), np.float32(0)), np.float32(6.0), np.nextafter(np.float32(6.0), np.float32(12))][k // 2 % 3]
            amides_k.center[k] = [8.0 * k, 50.0, 0.0]
            amides_k.center[k + 1] = [np.float32(8.0 * k) + d32, 50.0, 0.0]
            amides_k.normal[k] = amides_k.normal[k + 1] = [0.0, 0.0, 1.0]
        small = synth.cloud_featured(1_025, seed=34, bonds=False)
        small.xyz[:64] = (rings_k.center + np.array([0.0, 0.0, 6.0])).astype(np.float32)      # atoms exactly 6 A above the centroids
        _all_terms(eng, small, rings_k, amides_k, p, 'knife edge')
        # a NaN and an infinite centre: the reference's `distance > threshold` does not reject NaN
        rings_n, amides_n = synth.plane_set(40, 40, n_atoms=100, seed=35, n_residues=8)
        rings_n.center[3, 1] = np.nan
        rings_n.center[7, 0] = np.inf
        amides_n.center[5, 2] = np.nan
        _all_terms(eng, small, rings_n, amides_n, p, 'non-finite')


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


def test_all_terms_in_one_sequence(engine):
    """arp_planes_run_all == the four single-term runs == the oracle (configs[3] sizes)."""
    p = arp_params.make_params()
    engine.set_params(p)
    soa = synth.cloud_featured(30_000, seed=12, bonds=False)
    rings, amides = synth.plane_set(700, 4_000, 30_000, seed=13)
    engine.upload_atoms(soa)
    engine.upload_planes(rings, amides)
    got = engine.planes_all()
    util.assert_records_equal(got['ring_ring'], oracle.ring_ring(rings, p), 'all: ring-ring')
    util.assert_records_equal(got['atom_ring'], oracle.atom_ring(soa, rings, p), 'all: atom-ring')
    util.assert_records_equal(got['amide_amide'], oracle.amide_amide(amides, p), 'all: amide-amide')
    util.assert_records_equal(got['amide_ring'], oracle.amide_ring(amides, rings, p), 'all: amide-ring')
    util.assert_records_equal(engine.atom_ring(), got['atom_ring'], 'single run after run_all')
    util.assert_records_equal(engine.ring_ring(), got['ring_ring'], 'single run after run_all')


def test_plane_grid_edge_cases(engine):
    """Points far outside the atoms' bounding box, planes on cell borders, far-from-origin coordinates (float32 images
    of the float64 ring centroids), a tiny and a huge radius: the grid search must find what the double loop finds."""
    soa = synth.cloud_featured(3_000, seed=14, bonds=False)
    rings, amides = synth.plane_set(300, 900, 3_000, seed=15)
    rings.center[:20] += 500.0                       # outside every other point
    amides.center[:20] -= np.float32(300.0)
    rings.center[20:60] = np.round(rings.center[20:60] / 6.0) * 6.0     # on multiples of the default radius
    for shift in (0.0, 20000.0):
        soa2 = synth.cloud_featured(3_000, seed=14, bonds=False)
        soa2.xyz += np.float32(shift)
        r2 = type(rings)(rings.center + shift, rings.normal, rings.res_id, rings.flags, False)
        a2 = type(amides)((amides.center + np.float32(shift)).astype(np.float32), amides.normal, amides.res_id, amides.flags, True)
        for radius in (6.0, 0.7, 25.0):
            p = arp_params.make_params()
            p.ring_centroid_dist = p.amide_centroid_dist = p.met_sulphur_dist = radius
            engine.set_params(p)
            engine.upload_atoms(soa2)
            engine.upload_planes(r2, a2)
            got = engine.planes_all()
            what = f'shift {shift} radius {radius}'
            util.assert_records_equal(got['ring_ring'], oracle.ring_ring(r2, p), what + ' ring-ring')
            util.assert_records_equal(got['atom_ring'], oracle.atom_ring(soa2, r2, p), what + ' atom-ring')
            util.assert_records_equal(got['amide_amide'], oracle.amide_amide(a2, p), what + ' amide-amide')
            util.assert_records_equal(got['amide_ring'], oracle.amide_ring(a2, r2, p), what + ' amide-ring')
    engine.set_params(arp_params.make_params())


def test_non_finite_centres_take_the_double_loop(engine):
    """A NaN centre never compares `distance > threshold` true in the reference, so that plane pairs with every other
    one: the upload detects it and the plain double loops run (same records as the oracle)."""
    p = arp_params.make_params()
    engine.set_params(p)
    rings, amides = synth.plane_set(120, 200, n_atoms=2_000, seed=16, n_residues=40)
    rings.center[3, 1] = np.nan
    amides.center[5, 0] = np.inf
    engine.upload_planes(rings, amides)
    util.assert_records_equal(engine.ring_ring(), oracle.ring_ring(rings, p), 'ring-ring NaN centre')
    util.assert_records_equal(engine.amide_amide(), oracle.amide_amide(amides, p), 'amide-amide inf centre')
    util.assert_records_equal(engine.amide_ring(), oracle.amide_ring(amides, rings, p), 'amide-ring non-finite')


def test_plane_double_loop_knob(monkeypatch):
    """ARPEGGIO_NO_PLANE_GRID: the plain double loops give the same records (A/B knob and fallback path)."""
    from arpeggio_b200.engine import ContactEngine
    monkeypatch.setenv('ARPEGGIO_NO_PLANE_GRID', '1')
    p = arp_params.make_params()
    soa = synth.cloud_featured(5_000, seed=17, bonds=False)
    rings, amides = synth.plane_set(200, 800, 5_000, seed=18)
    with ContactEngine(0, p) as eng:
        eng.upload_atoms(soa)
        eng.upload_planes(rings, amides)
        got = eng.planes_all()
    util.assert_records_equal(got['ring_ring'], oracle.ring_ring(rings, p), 'knob ring-ring')
    util.assert_records_equal(got['atom_ring'], oracle.atom_ring(soa, rings, p), 'knob atom-ring')
    util.assert_records_equal(got['amide_amide'], oracle.amide_amide(amides, p), 'knob amide-amide')
    util.assert_records_equal(got['amide_ring'], oracle.amide_ring(amides, rings, p), 'knob amide-ring')
