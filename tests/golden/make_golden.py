#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference; the fixtures travel, this script's
inputs do not):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

How: BioPython / OpenBabel / gemmi are not installable here, so ``tests/mockbio.py`` registers
duck-typed stand-ins under their module names; then ``arpeggio.core.interactions`` is imported
UNMODIFIED from /root/reference and, for each seeded mini complex, the reference's

    InteractionComplex._initialize_atom_sift / _initialize_residue_sift
    InteractionComplex.run_arpeggio -> _make_selection, _calculate_atom_contacts,
        _calculate_ring_contacts, _calculate_group_contacts          (interactions.py:329-347)
    InteractionComplex.get_contacts                                  (interactions.py:172-212)

run on it with the live NumPy.  What is third-party and therefore restated, not executed:
Bio.PDB.NeighborSearch (brute force in double, index1 < index2) and the OpenBabel object model.

Each fixture <case>.npz holds the packed SoA input (made by the product packer
arpeggio_b200.packing.pack_complex), the reference's contact records as flat arrays, the
``get_contacts()`` JSON, and the recipe needed to rebuild the same mock complex without the
reference (tests on the GPU box).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import mockbio  # noqa: E402

mockbio.install_stubs()
sys.path.insert(0, '/root/reference')
from arpeggio.core import interactions as ref_interactions  # noqa: E402

from arpeggio_b200 import abi  # noqa: E402
from arpeggio_b200.packing import pack_complex  # noqa: E402

CASES = {
    # name: (build_complex kwargs, selection strings, cutoff, vdw_comp, include_sequence_adjacent)
    'ligand_site': (dict(seed=11, n_chains=2, n_res=22, n_waters=14), ['RESNAME:LIG'], 5.0, 0.1, False),
    'whole_structure': (dict(seed=12, n_chains=2, n_res=14, n_waters=10), [], 5.0, 0.1, False),
    'seq_adjacent': (dict(seed=13, n_chains=1, n_res=20, n_waters=8), ['/A//'], 5.0, 0.1, True),
    'wide_cutoff': (dict(seed=14, n_chains=2, n_res=12, n_waters=6), ['/A/3/', '/A/4/', 'RESNAME:LIG'], 6.5, 0.25, False),
    'degenerate': (dict(seed=15, n_chains=1, n_res=12, n_waters=6, degenerate=True), ['RESNAME:DEG', 'RESNAME:LIG'], 5.0, 0.1, False),
    'xbond_fault': (dict(seed=16, n_chains=1, n_res=10, n_waters=4, lone_xdonor=True), ['RESNAME:LIG'], 5.0, 0.1, False),
    'tight_cutoff': (dict(seed=21, n_chains=2, n_res=18, n_waters=10), ['RESNAME:LIG'], 3.5, 0.0, False),
    'big_site': (dict(seed=22, n_chains=3, n_res=40, n_waters=40), ['/A/5/', '/B/7/', 'RESNAME:LIG'], 5.0, 0.1, False),
    'no_explicit_h': (dict(seed=23, n_chains=2, n_res=16, n_waters=8, explicit_h=False), ['RESNAME:LIG'], 5.0, 0.1, False),
    'apo_adjacent': (dict(seed=24, n_chains=2, n_res=20, n_waters=12, ligand=False), [], 4.0, 0.2, True),
    # protein-like density (backbone in a 1.5x larger sphere) + 100 / 90 copies of the geometries the rare rules need
    # (C-Cl...O, C-H...Br-C, N+...O-, Zn with three O ligands, water clusters): xbond, halogen weak hbond, ionic,
    # metal complex and WATER_WATER get >= 50 positive records each from the reference's own code
    'rich_whole': (dict(seed=31, n_chains=2, n_res=30, n_waters=10, motifs=100, spread=1.5), [], 5.0, 0.1, False),
    'rich_site': (dict(seed=32, n_chains=2, n_res=24, n_waters=8, motifs=90, spread=1.4),
                  ['RESNAME:CLX', 'RESNAME:BRX', 'RESNAME:LYX', 'RESNAME:ZN', 'RESNAME:HOH', 'RESNAME:LIG'], 4.5, 0.15, False),
    # the pairs of search_all in a seeded shuffle instead of ascending (i, j) (stand-in for Bio.PDB's KD-tree order):
    # the reference's selection_plus = list(set(...)) then comes out in another order than an insertion in entity
    # order gives, and integer_sift (utils.py:233) is that of another last contact
    'kd_order': (dict(seed=33, n_chains=2, n_res=18, n_waters=10), ['RESNAME:LIG'], 5.0, 0.1, False),
}
PAIR_ORDER_SEED = {'kd_order': 5}


RESIDUE_PLANE_SIFTS = ('ring_ring_inter_integer_sift', 'ring_atom_inter_integer_sift', 'atom_ring_inter_integer_sift',
                       'mc_atom_ring_inter_integer_sift', 'sc_atom_ring_inter_integer_sift', 'amide_ring_inter_integer_sift',
                       'ring_amide_inter_integer_sift', 'amide_amide_inter_integer_sift')


def reference_complex(cx):
    """An InteractionComplex whose __init__ (file parsing) is skipped and whose fields are the mock's."""
    ic = object.__new__(ref_interactions.InteractionComplex)
    ic.__dict__.update(cx.__dict__)
    ic.params = ref_interactions.Parameters(vdw_comp_factor=0.1, interacting_threshold=5.0,
                                            has_hydrogens=True, ph=7.4)
    # fields _initialize_atom_sift reads (interactions.py:1804-1815); irrelevant to the contacts
    for a in ic.s_atoms:
        a.atomic_number = ic.ob_mol.GetAtomById(ic.bio_to_ob[a]).GetAtomicNum()
        a.bond_order = 1
        a.formal_charge = 0
        a.num_hydrogens = len(a.h_coords)
    ic._initialize_atom_sift()
    ic._initialize_residue_sift()           # resets is_polypeptide ...
    mockbio.flag_polypeptides(ic)           # ... which _handle_chains_residues_and_breaks then sets
    return ic


def mask_of(sifts, text):
    m = 0
    for k, v in enumerate(sifts):
        assert v in (0, 1)
        m |= int(v) << k
    return m | (abi.CLASS_NAMES.index(text) << abi.CLASS_SHIFT)


def ring_assignment(kwargs):
    """The reference's _assign_aromatic_rings_to_residues (interactions.py:1453-1492) on a fresh copy of the
    complex, plus one ring far from every atom (-> residue None)."""
    cx = mockbio.build_complex(**kwargs)
    ic = reference_complex(cx)
    rings = ic.biopython_str.rings
    far = max(rings) + 1 if rings else 0
    rings[far] = {'ring_id': far, 'center': np.array([900.0, -900.0, 900.0]), 'atoms': []}
    for r in rings.values():
        r.pop('residue', None)
        r.pop('residue_shortest_distance', None)
    for res in ic.biopython_str.get_residues():
        if hasattr(res, 'rings'):
            del res.rings
    ic._assign_aromatic_rings_to_residues()
    res_index = {id(r): k for k, r in enumerate(ic.biopython_str.get_residues())}
    keys = list(rings)
    out = dict(
        f4_xyz=np.array([a.coord for a in ic.s_atoms], dtype=np.float32).reshape(-1, 3),
        f4_atom_res=np.array([res_index[id(a.get_parent())] for a in ic.s_atoms], dtype=np.int32),
        f4_centers=np.array([rings[k]['center'] for k in keys], dtype=np.float64).reshape(-1, 3),
        f4_ring_res=np.array([res_index[id(rings[k]['residue'])] if rings[k]['residue'] is not None else -1 for k in keys], dtype=np.int32),
        f4_ring_dist=np.array([rings[k].get('residue_shortest_distance', 0.0) for k in keys], dtype=np.float64))
    assert out['f4_ring_res'][-1] == -1
    for k in keys:                                       # the bookkeeping on the residue side (:1488-1491)
        assert rings[k]['residue'] is None or k in rings[k]['residue'].rings
    return out


def compute_case(name, kwargs, selections, cutoff, vdw_comp, incl, verbose=True):
    """Run the reference on one seeded mock complex; returns the arrays of a fixture (nothing is written)."""
    cx = mockbio.build_complex(**kwargs)
    ic = reference_complex(cx)
    meta = dict(case=name, recipe=kwargs, selections=selections, cutoff=cutoff, vdw_comp=vdw_comp,
                include_sequence_adjacent=incl, numpy=np.__version__, raises=None,
                pair_order_seed=PAIR_ORDER_SEED.get(name))
    mockbio.NeighborSearch.pair_order_seed = PAIR_ORDER_SEED.get(name)
    try:
        ic.run_arpeggio(selections, cutoff, vdw_comp, incl)
    except AttributeError as err:           # utils.py:173 on a donor without single-bond neighbour
        meta['raises'] = 'AttributeError'
        if verbose:
            print(f'  {name}: reference raised AttributeError ({err})')
        # selection bookkeeping is complete at that point; contacts are not
        ic.atom_contacts = []
    mockbio.NeighborSearch.pair_order_seed = None
    packed = pack_complex(ic)
    idx = {id(a): i for i, a in enumerate(packed.atoms)}
    ring_idx = {k: i for i, k in enumerate(packed.ring_keys)}
    amide_idx = {k: i for i, k in enumerate(packed.amide_keys)}

    pairs = np.zeros(len(ic.atom_contacts), dtype=abi.PAIR_DTYPE)
    for k, c in enumerate(ic.atom_contacts):
        i, j = idx[id(c.bgn_atom)], idx[id(c.end_atom)]
        assert i < j, 'NeighborSearch reports index1 < index2'
        assert c.distance.dtype == np.float32
        pairs[k] = (i, j, mask_of(c.sifts, c.contact_type), c.distance)
    pairs = pairs[np.lexsort((pairs['j'], pairs['i']))]

    def plane_records(contacts, idx_a, idx_b, geom):
        out = np.zeros(len(contacts), dtype=abi.PLANE_PAIR_DTYPE)
        for k, c in enumerate(contacts):
            if geom:
                g = [abi.GEOM_NAMES.index(t) for t in c.contact_type]
                code = g[0] | ((g[1] if len(g) > 1 else 0xF) << 4)
                assert len(g) <= 2
            else:
                code = 0xFF
            code |= abi.CLASS_NAMES.index(c.text) << 8
            code |= int(c.bgn_res == c.end_res) << 11
            out[k] = (idx_a[c.bgn_id], idx_b[c.end_id], code, 0, np.float64(c.distance))
        return out

    rr = plane_records(ic.plane_plane_contacts, ring_idx, ring_idx, True)
    aa = plane_records(ic.group_group_contacts, amide_idx, amide_idx, False)
    ar = plane_records(ic.group_plane_contacts, amide_idx, ring_idx, False)
    # atom-plane: the record has no ring id; recover it from the ring's residue + atom names + distance
    ap = np.zeros(len(ic.atom_plane_contacts), dtype=abi.ATOM_PLANE_DTYPE)
    rings = ic.biopython_str.rings
    for k, c in enumerate(ic.atom_plane_contacts):
        cands = [key for key in packed.ring_keys
                 if rings[key]['residue'] == c.end_res
                 and sorted(a.get_id() for a in rings[key]['atoms']) == c.end_res_atoms
                 and np.linalg.norm(c.bgn_atom.coord - rings[key]['center']) == c.distance]
        assert len(cands) >= 1
        code = 0
        for lab in c.sifts:
            code |= 1 << abi.AP_NAMES.index(lab)
        code |= abi.CLASS_NAMES.index(c.text) << 8
        code |= int(c.end_res == c.bgn_atom.get_parent()) << 11
        ap[k] = (idx[id(c.bgn_atom)], ring_idx[cands[0]], code, 0, np.float64(c.distance))
    ap = ap[np.lexsort((ap['atom'], ap['ring']))]

    contacts_json = json.dumps(ic.get_contacts(), sort_keys=True) if meta['raises'] is None else '[]'

    # the pair loop's side effects on the atoms (utils.py:182-242, interactions.py:822-852), list order
    atom_sifts = np.zeros(len(packed.atoms), dtype=abi.ATOM_SIFT_DTYPE)
    for i, a in enumerate(packed.atoms):
        for c, suffix in enumerate(abi.SIFT_CATEGORIES):
            sift, integer = getattr(a, 'sift' + suffix), getattr(a, 'integer_sift' + suffix)
            fsift = getattr(a, 'actual_fsift' + suffix)
            assert len(sift) == 15 and len(integer) == 15 and list(fsift) == list(sift[5:])
            assert set(sift) <= {0, 1, True, False} and set(integer) <= {0, 1, 2}
            atom_sifts['sift'][i, c] = sum(int(bool(v)) << b for b, v in enumerate(sift))
            atom_sifts['integer_sift'][i, c] = sum(int(v) << (2 * b) for b, v in enumerate(integer))
            atom_sifts['hbonds'][i, c] = getattr(a, 'actual_hbonds' + suffix)
            atom_sifts['polars'][i, c] = getattr(a, 'actual_polars' + suffix)
    # per-residue counters of the plane loops (interactions.py:1040-1057, :1171-1176, :1290-1291, :1371-1373)
    res_sifts = {}
    for r in ic.biopython_str.get_residues():
        entry = {n: list(getattr(r, n)) for n in RESIDUE_PLANE_SIFTS}
        if any(any(v) for v in entry.values()):
            res_sifts[mockbio.residue_key(r)] = entry
    meta['residue_plane_sifts'] = res_sifts

    meta.update(
        selection_serials=[a.serial_number for a in ic.selection],
        selection_plus_serials=[a.serial_number for a in ic.selection_plus],
        selection_ring_ids=sorted(ic.selection_ring_ids), selection_plus_ring_ids=sorted(ic.selection_plus_ring_ids),
        selection_amide_ids=sorted(ic.selection_amide_ids), selection_plus_amide_ids=sorted(ic.selection_plus_amide_ids),
        n_atoms=len(packed.atoms), n_pairs=int(len(pairs)), n_ring_ring=int(len(rr)), n_atom_ring=int(len(ap)),
        n_amide_amide=int(len(aa)), n_amide_ring=int(len(ar)))
    s = packed.soa
    arrays = dict(
        xyz=s.xyz, feat=s.feat, res_id=s.res_id, rad_class=s.rad_class, vdw=s.vdw, cov=s.cov,
        res_prev=s.res_prev, res_next=s.res_next, res_flags=s.res_flags, bond_off=s.bond_off,
        bond_nbr=s.bond_nbr, h_off=s.h_off, h_xyz=s.h_xyz,
        xnbr_xyz=s.xnbr_xyz if s.xnbr_xyz is not None else np.zeros((0, 3), np.float32),
        ring_center=packed.rings.center, ring_normal=packed.rings.normal, ring_res=packed.rings.res_id,
        ring_flags=packed.rings.flags,
        amide_center=packed.amides.center, amide_normal=packed.amides.normal, amide_res=packed.amides.res_id,
        amide_flags=packed.amides.flags,
        exp_atom_sifts=atom_sifts, exp_pairs=pairs, exp_ring_ring=rr, exp_atom_ring=ap, exp_amide_amide=aa, exp_amide_ring=ar,
        meta=np.array(json.dumps(meta)), contacts_json=np.array(contacts_json), **ring_assignment(kwargs))
    return arrays


def run_case(name, kwargs, selections, cutoff, vdw_comp, incl):
    arrays = compute_case(name, kwargs, selections, cutoff, vdw_comp, incl)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    pairs, rr, ap, aa, ar = (arrays[k] for k in ('exp_pairs', 'exp_ring_ring', 'exp_atom_ring', 'exp_amide_amide', 'exp_amide_ring'))
    n_atoms = arrays['xyz'].shape[0]
    print(f'  {name}: N={n_atoms} pairs={len(pairs)} ring-ring={len(rr)} atom-ring={len(ap)} '
          f'amide-amide={len(aa)} amide-ring={len(ar)}')
    from collections import Counter
    bits = Counter()
    for m in pairs['mask']:
        for b in range(15):
            if m >> b & 1:
                bits[abi.SIFT_NAMES[b]] += 1
    print('     bits:', dict(bits))
    hal = (arrays['feat'] & abi.F_IS_HALOGEN) != 0
    n_hal_weak = int(np.count_nonzero((pairs['mask'] >> 6 & 1).astype(bool) & (hal[pairs['i']] | hal[pairs['j']])))
    print('     weak_hbond records with a halogen (is_halogen_weak_hbond):', n_hal_weak)
    cls = Counter(abi.CLASS_NAMES[(int(m) >> abi.CLASS_SHIFT) & abi.CLASS_MASK] for m in pairs['mask'])
    print('     classes:', dict(cls))


if __name__ == '__main__':
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(name, *spec)
