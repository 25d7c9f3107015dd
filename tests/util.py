"""Shared helpers of the test-suite: golden fixture loading and record canonicalisation."""
import glob
import json
import os

import numpy as np

from arpeggio_b200 import abi, params as arp_params
from arpeggio_b200.soa import AtomSoA, PlaneSoA

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.name = name
        self.meta = json.loads(str(z['meta']))
        # the reference saw the pairs of search_all in a seeded shuffle (stand-in for Bio.PDB's KD-tree order): everything
        # order-dependent (integer_sift, utils.py:233) is then NOT reproducible from the (i, j)-sorted stream
        self.shuffled_pairs = self.meta.get('pair_order_seed') is not None
        self.contacts_json = json.loads(str(z['contacts_json']))
        xn = z['xnbr_xyz']
        self.soa = AtomSoA(xyz=z['xyz'], feat=z['feat'], res_id=z['res_id'], rad_class=z['rad_class'], vdw=z['vdw'],
                           cov=z['cov'], res_prev=z['res_prev'], res_next=z['res_next'], res_flags=z['res_flags'],
                           bond_off=z['bond_off'], bond_nbr=z['bond_nbr'], h_off=z['h_off'], h_xyz=z['h_xyz'],
                           xnbr_xyz=xn if xn.shape[0] else None)
        self.rings = PlaneSoA(z['ring_center'], z['ring_normal'], z['ring_res'], z['ring_flags'], False)
        self.amides = PlaneSoA(z['amide_center'], z['amide_normal'], z['amide_res'], z['amide_flags'], True)
        self.exp_atom_sifts = z['exp_atom_sifts']
        self.f4 = {k: z['f4_' + k] for k in ('xyz', 'atom_res', 'centers', 'ring_res', 'ring_dist')}
        self.exp_pairs = z['exp_pairs']
        self.exp_ring_ring = z['exp_ring_ring']
        self.exp_atom_ring = z['exp_atom_ring']
        self.exp_amide_amide = z['exp_amide_amide']
        self.exp_amide_ring = z['exp_amide_ring']
        m = self.meta
        self.params = arp_params.make_params(m['cutoff'], m['vdw_comp'], m['include_sequence_adjacent'])


def sort_pairs(rec):
    return rec[np.lexsort((rec['j'], rec['i']))]


def sort_planes(rec, keys=('a', 'b')):
    return rec[np.lexsort((rec[keys[1]], rec[keys[0]]))]


def assert_records_equal(got, exp, what, dist_bits=True):
    """Bit-exact comparison of two record arrays (same order expected)."""
    assert got.shape == exp.shape, f'{what}: {got.shape[0]} records, expected {exp.shape[0]}'
    for f in got.dtype.names:
        if f == '_pad':
            continue
        g, e = got[f], exp[f]
        if f == 'dist' and dist_bits:
            it = np.uint32 if g.dtype == np.float32 else np.uint64
            both_nan = np.isnan(g) & np.isnan(e)         # a NaN is a NaN: sign and payload are not part of the contract
            g, e = np.where(both_nan, it(0), g.view(it)), np.where(both_nan, it(0), e.view(it))
        bad = np.nonzero(g != e)[0]
        assert bad.size == 0, f'{what}: field {f} differs at {bad[:5]}: got {got[bad[:5]]}, expected {exp[bad[:5]]}'


def describe_mask(m):
    return [abi.SIFT_NAMES[b] for b in range(15) if m >> b & 1] + [abi.CLASS_NAMES[(m >> 16) & 7]]


def assert_atom_sifts_equal(got, g, what=''):
    """arp_atom_sift arrays against a fixture: every field; for a fixture whose reference run saw shuffled pairs the
    order-dependent integer_sift is exempt -- and must indeed differ, or the fixture pins nothing."""
    for f in got.dtype.names:
        if f == 'integer_sift' and g.shuffled_pairs:
            assert not np.array_equal(got[f], g.exp_atom_sifts[f]), 'integer_sift did not depend on the pair order'
            continue
        assert np.array_equal(got[f], g.exp_atom_sifts[f]), f'{what} {f}'
