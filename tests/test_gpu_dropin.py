"""The drop-in surface on the GPU: the reference's run sequence (run_arpeggio -> _calculate_atom_contacts,
_calculate_ring_contacts, _calculate_group_contacts -> get_contacts) over the CUDA mixin must give the JSON
the real reference produced for the same complex (tests/golden/*.npz, `contacts_json`)."""
import json

import numpy as np
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
    # order of the sorted record stream: the lists agree element by element (not for the fixture whose reference
    # run saw the pairs in a shuffled, KD-tree-like order)
    if not g.shuffled_pairs:
        assert got == exp
    kinds = {e['type'] for e in got}
    assert 'atom-atom' in kinds


@pytest.mark.parametrize('case', [c for c in util.golden_cases() if c != 'xbond_fault'])
def test_sift_side_effects_match_reference(engine, case):
    """SURVEY 8 f3 through the drop-in: after run_arpeggio the atoms carry the attributes the reference's pair loop
    leaves (sift*, integer_sift*, actual_fsift*, actual_hbonds*, actual_polars*) and the residues the counters
    of the plane loops, value for value."""
    import mockbio
    from arpeggio_b200 import abi
    g = util.Golden(case)
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    host.cuda_integer_sifts = True      # opt in: the fixture's mock search yields (i, j)-sorted pairs, the order evaluated here
    m = g.meta
    host.run_arpeggio(m['cutoff'], m['vdw_comp'], m['include_sequence_adjacent'])
    exp = g.exp_atom_sifts
    for i, a in enumerate(host.selection_plus):
        for c, suffix in enumerate(abi.SIFT_CATEGORIES):
            sift = [int(exp['sift'][i, c]) >> b & 1 for b in range(15)]
            assert a.__dict__['sift' + suffix] == sift
            assert a.__dict__['actual_fsift' + suffix] == sift[5:]
            if not g.shuffled_pairs:        # order-dependent (utils.py:233): only comparable when the reference saw sorted pairs
                assert a.__dict__['integer_sift' + suffix] == [int(exp['integer_sift'][i, c]) >> 2 * b & 3 for b in range(15)]
            assert a.__dict__['actual_hbonds' + suffix] == int(exp['hbonds'][i, c])
            assert a.__dict__['actual_polars' + suffix] == int(exp['polars'][i, c])
    want = m['residue_plane_sifts']
    seen = 0
    for r in host.biopython_str.get_residues():
        e = want.get(mockbio.residue_key(r))
        for name in ('ring_ring_inter_integer_sift', 'ring_atom_inter_integer_sift', 'atom_ring_inter_integer_sift',
                     'mc_atom_ring_inter_integer_sift', 'sc_atom_ring_inter_integer_sift', 'amide_ring_inter_integer_sift',
                     'ring_amide_inter_integer_sift', 'amide_amide_inter_integer_sift'):
            got = getattr(r, name)
            assert got == (e[name] if e else [0] * len(got)), (mockbio.residue_key(r), name)
        seen += e is not None
    assert seen == len(want)


@pytest.mark.parametrize('case', [c for c in util.golden_cases() if c != 'xbond_fault'])
@pytest.mark.parametrize('lazy', [False, True])
def test_contacts_json_text_is_the_reference_dump(engine, case, lazy, tmp_path):
    """the file process_protein_cli.py:187-188 writes, byte for byte, from the C emitter; and the lazy
    atom_contacts sequence behaves like the reference's list."""
    g = util.Golden(case)
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    host.cuda_lazy_contacts = lazy
    m = g.meta
    host.run_arpeggio(m['cutoff'], m['vdw_comp'], m['include_sequence_adjacent'])
    want = json.dumps(g.contacts_json, indent=4, sort_keys=True)
    if g.shuffled_pairs:
        # the reference listed its atom-atom contacts in the (shuffled) order of its pair search; the drop-in lists them
        # by (bgn, end): the same entries, and the C emitter's text is the dump of the drop-in's own list
        canon = lambda entries: sorted(json.dumps(e, sort_keys=True) for e in entries)
        assert canon(json.loads(host.contacts_json_text())) == canon(g.contacts_json)
        want = json.dumps(host.get_contacts(), indent=4, sort_keys=True)
    assert host.contacts_json_text() == want
    assert json.dumps(host.get_contacts(), indent=4, sort_keys=True) == want
    host.write_contacts_json(tmp_path / 'out.json')
    assert (tmp_path / 'out.json').read_text() == want
    ac = host.atom_contacts
    assert len(ac) == g.exp_pairs.shape[0]
    if lazy and len(ac):
        assert ac[0] == list(ac)[0] == ac[0:1][0] and ac[-1] == list(ac)[-1]
        inter = list(filter(lambda c: c.contact_type == 'INTER', ac))
        assert len(inter) == int(np.count_nonzero((g.exp_pairs['mask'] >> 16 & 7) == 2))


@pytest.mark.parametrize('case', util.golden_cases())
def test_ring_assignment_through_the_dropin(engine, case):
    """_assign_aromatic_rings_to_residues of the mixin leaves the rings and residues as the
    reference's function did (tests/golden: f4_*), including a ring with no atom within 3 A."""
    g = util.Golden(case)
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    rings = host.biopython_str.rings
    far = max(rings) + 1 if rings else 0
    rings[far] = {'ring_id': far, 'center': np.array([900.0, -900.0, 900.0]), 'atoms': []}
    for r in rings.values():
        r.pop('residue', None)
        r.pop('residue_shortest_distance', None)
    residues = list(host.biopython_str.get_residues())
    for res in residues:
        res.__dict__.pop('rings', None)
    host._assign_aromatic_rings_to_residues()
    index = {id(r): k for k, r in enumerate(residues)}
    got_res = [index[id(rings[k]['residue'])] if rings[k]['residue'] is not None else -1 for k in rings]
    assert got_res == g.f4['ring_res'].tolist()
    got_dist = np.array([rings[k].get('residue_shortest_distance', 0.0) for k in rings])
    assert np.array_equal(got_dist.view(np.uint64), g.f4['ring_dist'].view(np.uint64))
    for k in rings:
        assert rings[k]['residue'] is None or k in rings[k]['residue'].rings
    assert isinstance(host.ns, type(host.ns)) and len(host.ns.atom_list) == len(host.s_atoms)


@pytest.mark.parametrize('case', util.golden_cases())
def test_make_selection_through_the_dropin(engine, case):
    """the binding-site expansion of _make_selection (interactions.py:1420-1424) from the GPU flags gives
    the reference's selection_plus (as a set: its list order is a set's iteration order there too) and id sets."""
    g = util.Golden(case)
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    m = g.meta
    by_serial = {a.serial_number: a for a in host.s_atoms}
    want_sel = [by_serial[s] for s in m['selection_serials']]
    host._cuda_parse_selection = lambda selections, entity: list(want_sel)
    for name in ('selection', 'selection_plus', 'selection_ring_ids', 'selection_plus_ring_ids', 'selection_amide_ids',
                 'selection_plus_amide_ids'):
        host.__dict__.pop(name, None)
    host._make_selection(m['selections'] or ['x'])
    assert [a.serial_number for a in host.selection] == m['selection_serials']
    assert len(host.selection_plus) == len(m['selection_plus_serials'])
    assert {a.serial_number for a in host.selection_plus} == set(m['selection_plus_serials'])
    assert sorted(host.selection_ring_ids) == m['selection_ring_ids']
    assert sorted(host.selection_plus_ring_ids) == m['selection_plus_ring_ids']
    assert sorted(host.selection_amide_ids) == m['selection_amide_ids']
    assert sorted(host.selection_plus_amide_ids) == m['selection_plus_amide_ids']
    assert host.selection_plus_residues == {a.get_parent() for a in host.selection_plus}
    assert len(host.ns.atom_list) == len(host.selection_plus)
    with pytest.raises(AttributeError):
        host._cuda_parse_selection = lambda selections, entity: []
        host._make_selection(['nothing'])


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


def test_integer_sift_is_opt_in(engine):
    """atom.integer_sift* depends on the reference's KD-tree loop order (utils.py:233): without the opt-in the drop-in
    removes the attributes so that a consumer fails loudly instead of reading sorted-order values."""
    g = util.Golden('ligand_site')
    host = mock_host.host_from_golden(g)
    host.cuda_engine = engine
    m = g.meta
    host.run_arpeggio(m['cutoff'], m['vdw_comp'], m['include_sequence_adjacent'])
    a = host.selection_plus[0]
    assert 'sift' in a.__dict__ and 'actual_hbonds' in a.__dict__
    assert not any(k.startswith('integer_sift') for k in a.__dict__)
