"""The CPU oracle against the reference's own code run LIVE on freshly seeded mock complexes (only where
/root/reference exists, i.e. in the build container; the committed fixtures of tests/golden/ carry the same
comparison to the GPU box).  A sweep over seeds, sizes, cutoffs and VdW compensations, so that the oracle is
pinned on far more geometry than the ten committed cases."""
import os
import sys

import numpy as np
import pytest

import util
from arpeggio_b200 import abi, params as arp_params
from arpeggio_b200.soa import AtomSoA, PlaneSoA
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/arpeggio'), reason='the reference tree is not on this box')


@pytest.fixture(scope='module')
def make_golden():
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import make_golden as mg          # installs the Bio / openbabel stand-ins and imports the reference unmodified
    return mg


SWEEP = [
    # (complex recipe, selections, cutoff, vdw_comp, include_sequence_adjacent)
    (dict(seed=100 + k, n_chains=1 + k % 3, n_res=10 + 3 * (k % 5), n_waters=4 + 2 * (k % 4), explicit_h=k % 4 != 3),
     [[], ['RESNAME:LIG'], ['/A/3/', 'RESNAME:LIG'], ['/A//']][k % 4],
     [5.0, 4.0, 6.0, 3.0, 7.5][k % 5], [0.1, 0.0, 0.25, 0.4][k % 4], k % 3 == 0)
    for k in range(16)
]


@pytest.mark.parametrize('k', range(len(SWEEP)))
def test_oracle_equals_the_running_reference(make_golden, k):
    recipe, selections, cutoff, comp, adjacent = SWEEP[k]
    z = make_golden.compute_case(f'sweep{k}', recipe, selections, cutoff, comp, adjacent, verbose=False)
    import json
    meta = json.loads(str(z['meta']))
    if meta['raises']:
        pytest.skip('the reference raised on this complex (is_xbond without a neighbour)')
    xn = z['xnbr_xyz']
    soa = AtomSoA(xyz=z['xyz'], feat=z['feat'], res_id=z['res_id'], rad_class=z['rad_class'], vdw=z['vdw'], cov=z['cov'],
                  res_prev=z['res_prev'], res_next=z['res_next'], res_flags=z['res_flags'], bond_off=z['bond_off'],
                  bond_nbr=z['bond_nbr'], h_off=z['h_off'], h_xyz=z['h_xyz'], xnbr_xyz=xn if xn.shape[0] else None)
    rings = PlaneSoA(z['ring_center'], z['ring_normal'], z['ring_res'], z['ring_flags'], False)
    amides = PlaneSoA(z['amide_center'], z['amide_normal'], z['amide_res'], z['amide_flags'], True)
    p = arp_params.make_params(cutoff, comp, adjacent)
    what = f'sweep {k}: {recipe} {selections} cutoff {cutoff} comp {comp} adjacent {adjacent}'
    got = oracle.pairs(soa, p)
    assert not np.any(got['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR))
    util.assert_records_equal(got, z['exp_pairs'], what + ' atom-atom')
    util.assert_records_equal(oracle.ring_ring(rings, p), z['exp_ring_ring'], what + ' ring-ring')
    util.assert_records_equal(oracle.atom_ring(soa, rings, p), z['exp_atom_ring'], what + ' atom-ring')
    util.assert_records_equal(oracle.amide_amide(amides, p), z['exp_amide_amide'], what + ' amide-amide')
    util.assert_records_equal(oracle.amide_ring(amides, rings, p), z['exp_amide_ring'], what + ' amide-ring')
    sifts = oracle.atom_sifts(z['exp_pairs'], soa.n_atoms)
    for f in sifts.dtype.names:
        assert np.array_equal(sifts[f], z['exp_atom_sifts'][f]), what + ' per-atom ' + f


def test_default_thresholds_are_the_references_config():
    """params.DEFAULT_CONTACT_TYPES / DEFAULT_DIST_MAX / DEFAULT_H_VDW restate config.py:592-660 and :23-25; here they
    are compared with the reference's own module, key by key (the drop-in reads the live config anyway)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_config', '/root/reference/arpeggio/core/config.py')
    cfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg)
    ours = arp_params.DEFAULT_CONTACT_TYPES
    for kind, entries in ours.items():
        assert kind in cfg.CONTACT_TYPES, kind
        for key, value in entries.items():
            assert cfg.CONTACT_TYPES[kind][key] == value, (kind, key, cfg.CONTACT_TYPES[kind][key], value)
    assert cfg.CONTACT_TYPES_DIST_MAX == arp_params.DEFAULT_DIST_MAX
    assert cfg.VDW_RADII['H'] == arp_params.DEFAULT_H_VDW
