"""BASELINE configs[0]: `1tqn_h.cif -s /A/508/` through the UNMODIFIED reference and through the drop-in, side by
side, plus the four example records of the reference's README (README.md:127-241).

Needs what this image does not have -- BioPython, OpenBabel, gemmi and the structure file -- and skips itself
without them.  Where to put them: the reference tree on PYTHONPATH (or pip-installed), the file under
tests/data/1tqn_h.cif, baseline/_ref/1tqn_h.cif or $ARPEGGIO_1TQN."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get('ARPEGGIO_1TQN', ''), os.path.join(HERE, 'data', '1tqn_h.cif'),
              os.path.join(os.path.dirname(HERE), 'baseline', '_ref', '1tqn_h.cif')]


def _structure_file():
    for p in CANDIDATES:
        if p and os.path.isfile(p):
            return p
    return None


def _canonical(contacts):
    """get_contacts() entries in a canonical order (the reference's own order is KD-tree traversal order)."""
    return sorted(json.dumps(c, sort_keys=True) for c in contacts)


def test_1tqn_heme_site_json_identical_to_the_reference():
    for mod in ('Bio', 'openbabel', 'gemmi'):
        pytest.importorskip(mod)
    path = _structure_file()
    if path is None:
        pytest.skip('1tqn_h.cif not found (tests/data/, baseline/_ref/ or $ARPEGGIO_1TQN)')
    from arpeggio.core import InteractionComplex
    from arpeggio_b200.dropin import cuda_interaction_complex

    def run(cls):
        ic = cls(path, 0.1, 5.0, 7.4)
        ic.structure_checks()
        ic.initialize()
        ic.run_arpeggio(['/A/508/'], 5.0, 0.1, False)
        return ic.get_contacts()

    ref = run(InteractionComplex)
    ours = run(cuda_interaction_complex())
    assert _canonical(ours) == _canonical(ref)

    # the README's example records (README.md:127-241)
    def find(kind, pred):
        return [c for c in ours if c['type'] == kind and pred(c)]

    def atom(c, side, res, name):
        return c[side]['auth_seq_id'] == res and c[side]['auth_atom_id'] == name

    hits = find('atom-atom', lambda c: (atom(c, 'bgn', 313, 'CB') and atom(c, 'end', 508, 'CBB')) or
                (atom(c, 'end', 313, 'CB') and atom(c, 'bgn', 508, 'CBB')))
    assert hits and sorted(hits[0]['contact']) == ['hydrophobic', 'proximal'] and hits[0]['distance'] == 4.02
    assert hits[0]['interacting_entities'] == 'INTER'
    assert find('atom-plane', lambda c: c['contact'] == ['DONORPI'] and c['distance'] == 3.9)
    assert find('plane-plane', lambda c: c['contact'] == ['FT', 'ET'] and c['distance'] == 4.72)
    assert find('group-group', lambda c: c['contact'] == ['AMIDEAMIDE'] and c['distance'] == 4.29)
