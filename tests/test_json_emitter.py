"""the C emitter of the atom-atom contact JSON (csrc/arp_json.cu, host code -- runs without a GPU)
must reproduce json.dump(get_contacts(), indent=4, sort_keys=True) of the reference byte for byte."""
import json

import numpy as np
import pytest

import mock_host
import util
from arpeggio_b200 import abi, jsonout


def _fragments(host):
    frags = []
    for a in host.selection_plus:
        d = mock_host.make_pymol_json(a)
        d['label_comp_type'] = host.component_types[mock_host.get_residue_name(a)]
        frags.append(jsonout.atom_fragment(d))
    return frags


@pytest.mark.parametrize('case', [c for c in util.golden_cases() if c != 'xbond_fault'])
@pytest.mark.parametrize('threads', [1, 5])
def test_atom_atom_text_matches_reference_dump(case, threads):
    g = util.Golden(case)
    host = mock_host.host_from_golden(g)
    atom_atom = [e for e in g.contacts_json if e['type'] == 'atom-atom']
    others = [e for e in g.contacts_json if e['type'] != 'atom-atom']
    assert len(atom_atom) == g.exp_pairs.shape[0]
    body = jsonout.pairs_json(g.exp_pairs, _fragments(host), threads=threads)
    if g.shuffled_pairs:            # the reference listed the contacts in its (shuffled) pair order: same entries, other order
        canon = lambda entries: sorted(json.dumps(e, sort_keys=True) for e in entries)
        assert canon(json.loads('[' + bytes(body).decode() + ']')) == canon(atom_atom)
        return
    want = json.dumps(atom_atom, indent=4, sort_keys=True)
    assert '[\n' + bytes(body).decode() + '\n]' == want
    # the whole file, as process_protein_cli.py:187-188 writes it
    assert jsonout.splice(body, others) == json.dumps(g.contacts_json, indent=4, sort_keys=True)


def test_fragment_layout_matches_json_module():
    for d in ({}, {'b': 'x"y', 'a': 3, 'c': ' ', 'd': None, 'e': True, 'f': 1.5}, {'k': [1, 2], 'a': 'z'}, {'n': {'m': 1}}):
        entry = {'bgn': d}
        want = json.dumps([entry], indent=4, sort_keys=True)
        assert '[\n    {\n        "bgn": ' + jsonout.atom_fragment(d).decode() + '\n    }\n]' == want


def test_write_spliced(tmp_path):
    other = [{'type': 'plane-plane', 'contact': ['FF'], 'distance': 4.5}]
    rec = np.zeros(2, abi.PAIR_DTYPE)
    rec['j'] = 1
    rec['mask'] = 1 << 4
    body = jsonout.pairs_json(rec, [b'{}', b'{}'])
    for b_, o_ in ((body, other), (body, []), (np.zeros(0, np.uint8), other), (np.zeros(0, np.uint8), [])):
        with open(tmp_path / 'x.json', 'wb') as fp:
            jsonout.write_spliced(fp, b_, o_)
        assert (tmp_path / 'x.json').read_text() == jsonout.splice(b_, o_)
        json.loads((tmp_path / 'x.json').read_text())


def test_empty_and_splice_edges():
    assert bytes(jsonout.pairs_json(np.zeros(0, abi.PAIR_DTYPE), [])) == b''
    assert jsonout.splice(b'', []) == json.dumps([], indent=4, sort_keys=True) == '[]'
    other = [{'type': 'plane-plane', 'contact': ['FF'], 'distance': 4.5}]
    assert jsonout.splice(b'', other) == json.dumps(other, indent=4, sort_keys=True)


def test_distance_text_is_python_repr_of_numpy_round():
    """distance = round(np.float64(dist), 2) printed by float.__repr__ (interactions.py:190 + json encoder)."""
    rng = np.random.default_rng(7)
    special = [0.0, -0.0, 0.005, 0.015, 0.025, 1.005, 2.675, 2.665, 4.0, 4.5, 4.02, 3.9, 1e-5, 1e-3, 99.995, 100.0, 123456.78,
               8388608.0, 16777216.0, 1e13, 3e14, 1e15, 1e16, 1.2345e22, 3.4e38, 1e-30, 1e-45, np.nan, np.inf, -np.inf, -3.14159]
    vals = np.concatenate([np.array(special, np.float32), rng.uniform(0, 8, 20000).astype(np.float32),
                           (rng.integers(0, 800, 5000) / 100 + rng.choice([0.005, -0.005, 0.0], 5000)).astype(np.float32),
                           np.exp(rng.uniform(-20, 40, 3000)).astype(np.float32)])
    rec = np.zeros(vals.shape[0], abi.PAIR_DTYPE)
    rec['j'] = 1
    rec['mask'] = 1 << 4 | 2 << 16
    rec['dist'] = vals
    frag = [b'{}', b'{}']
    text = bytes(jsonout.pairs_json(rec, frag, threads=3)).decode()
    got = [ln.split(': ', 1)[1].rstrip(',') for ln in text.split('\n') if ln.startswith('        "distance"')]
    assert len(got) == vals.shape[0]
    want = [json.dumps(round(np.float64(v), 2)) for v in vals]
    bad = [(float(v), g_, w) for v, g_, w in zip(vals, got, want) if g_ != w]
    assert not bad, bad[:10]


def test_rejects_out_of_range_atoms():
    rec = np.zeros(1, abi.PAIR_DTYPE)
    rec['j'] = 5
    with pytest.raises(RuntimeError):
        jsonout.pairs_json(rec, [b'{}', b'{}'])


def test_every_mask_and_class():
    """All 2^15 SIFt words and all entity classes, against the reference's list comprehension."""
    masks = np.arange(1 << 15, dtype=np.uint32)
    rec = np.zeros(masks.shape[0], abi.PAIR_DTYPE)
    rec['j'] = 1
    rec['mask'] = masks | ((masks % 6) << 16)
    rec['dist'] = 3.25
    frag = [b'{\n            "auth_atom_id": "C1\\""\n        }', b'{}']
    got = json.loads('[' + bytes(jsonout.pairs_json(rec, frag)).decode() + ']')
    for m, e in zip(masks.tolist(), got):
        assert e['contact'] == [n for b, n in enumerate(abi.SIFT_NAMES) if m >> b & 1]
        assert e['interacting_entities'] == abi.CLASS_NAMES[m % 6]
        assert e['bgn'] == {'auth_atom_id': 'C1"'} and e['end'] == {} and e['distance'] == 3.25 and e['type'] == 'atom-atom'
