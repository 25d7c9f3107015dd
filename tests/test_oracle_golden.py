"""The CPU oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  This is what pins the oracle; the CUDA path is then
compared with the oracle (tests/test_gpu_*.py)."""
import numpy as np
import pytest

import util
from arpeggio_b200 import abi
from oracle import oracle

CASES = util.golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize('case', CASES)
def test_pairs_match_reference(case):
    g = util.Golden(case)
    got = oracle.pairs(g.soa, g.params)
    if g.meta['raises'] == 'AttributeError':
        # utils.py:173: the reference dies on an xbond donor without a single-bond neighbour;
        # the oracle reports the pair that would have raised
        assert np.any(got['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR))
        return
    assert not np.any(got['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR))
    util.assert_records_equal(got, g.exp_pairs, f'{case} atom-atom')


@pytest.mark.parametrize('case', CASES)
def test_pairs_bruteforce_equals_grid(case):
    g = util.Golden(case)
    util.assert_records_equal(oracle.pairs(g.soa, g.params, bruteforce=True), oracle.pairs(g.soa, g.params),
                              f'{case} grid vs brute force')


@pytest.mark.parametrize('case', [c for c in CASES if c != 'xbond_fault'])
def test_planes_match_reference(case):
    g = util.Golden(case)
    util.assert_records_equal(oracle.ring_ring(g.rings, g.params), g.exp_ring_ring, f'{case} ring-ring')
    util.assert_records_equal(oracle.atom_ring(g.soa, g.rings, g.params), g.exp_atom_ring, f'{case} atom-ring')
    util.assert_records_equal(oracle.amide_amide(g.amides, g.params), g.exp_amide_amide, f'{case} amide-amide')
    util.assert_records_equal(oracle.amide_ring(g.amides, g.rings, g.params), g.exp_amide_ring, f'{case} amide-ring')


@pytest.mark.parametrize('case', [c for c in CASES if c != 'xbond_fault'])
def test_atom_sifts_match_reference(case):
    """The per-atom side effects of the reference's pair loop (atom.sift*, integer_sift*, actual_hbonds*,
    actual_polars*; utils.py:182-242, interactions.py:822-852) replayed by the oracle over the records."""
    g = util.Golden(case)
    got = oracle.atom_sifts(g.exp_pairs, g.soa.n_atoms)
    util.assert_atom_sifts_equal(got, g, case)
    assert (got['integer_sift'] != 0).any() and (got['hbonds'][:, 0] > 0).any()
    # the integer SIFt is not the plain count-capped OR: the last contact decides (utils.py:233)
    two = (got['integer_sift'][:, 0][:, None] >> (2 * np.arange(15)) & 3) == 2
    assert two.any()


@pytest.mark.parametrize('case', CASES)
def test_ring_assignment_matches_reference(case):
    """_assign_aromatic_rings_to_residues (interactions.py:1453-1492): residue of the closest atom within 3 A of
    every ring centroid and that distance, float64 bits; rings with no atom nearby get None."""
    g = util.Golden(case)
    f = g.f4
    atom, dist = oracle.ring_nearest_atom(f['xyz'], f['centers'], 3.0, g.params)
    res = np.where(atom >= 0, f['atom_res'][np.maximum(atom, 0)], -1)
    assert np.array_equal(res, f['ring_res'])
    assert np.array_equal(dist.view(np.uint64), f['ring_dist'].view(np.uint64))
    assert (atom[-1] == -1) and (atom >= 0).any()
