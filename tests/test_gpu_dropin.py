"""The drop-in surface on the GPU: the reference's run sequence (run_arpeggio -> _calculate_atom_contacts,
_calculate_ring_contacts, _calculate_group_contacts -> get_contacts) over the CUDA mixin must give the JSON
the real reference produced for the same complex (tests/golden/*.npz, `contacts_json`)."""
import json

import pytest

import mock_host
import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', util.golden_cases())
def test_get_contacts_json_matches_reference(engine, case):
    g = util.Golden(case)
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    m = g.meta
    if m['raises'] == 'AttributeError':
        with pytest.raises(AttributeError):          # utils.py:173 in the reference
            host.run_arpeggio(m['cutoff'], m['vdw_comp'], m['include_sequence_adjacent'])
        return
    host.run_arpeggio(m['cutoff'], m['vdw_comp'], m['include_sequence_adjacent'])
    got = json.loads(json.dumps(host.get_contacts(), sort_keys=True))
    exp = g.contacts_json
    assert len(got) == len(exp)
    canon = lambda entries: sorted(json.dumps(e, sort_keys=True) for e in entries)
    assert canon(got) == canon(exp), 'contact JSON differs from the reference as a multiset'
    # the mock NeighborSearch of the fixture generator emits pairs by ascending (i, j), which is also the
    # order of the sorted record stream: the lists agree element by element
    assert got == exp
    kinds = {e['type'] for e in got}
    assert 'atom-atom' in kinds


def test_record_types_and_dtypes(engine):
    import numpy as np
    g = util.Golden('ligand_site')
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    host.run_arpeggio(5.0, 0.1, False)
    c = host.atom_contacts[0]
    assert type(c).__name__ == 'AtomAtomContact' and c._fields == ('bgn_atom', 'end_atom', 'sifts', 'contact_type', 'distance')
    assert isinstance(c.distance, np.float32) and len(c.sifts) == 15 and set(c.sifts) <= {0, 1}
    if host.plane_plane_contacts:
        p = host.plane_plane_contacts[0]
        assert p._fields[:3] == ('bgn_id', 'bgn_res', 'bgn_res_atoms') and isinstance(p.contact_type, list)
    if host.group_group_contacts:
        assert isinstance(host.group_group_contacts[0].distance, np.float32)
