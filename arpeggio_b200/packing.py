"""Host-side packer: an initialised ``InteractionComplex`` -> flat arrays.

Reads exactly the attributes the reference's contact loops read from the
BioPython / OpenBabel objects and lays them out as the SoA of
``include/arpeggio_cuda.h``.  Nothing here computes contacts.

Reference reads being mirrored (arpeggio/core/...):
  atom order            ``self.selection_plus`` list, interactions.py:1426, :1442
  coordinates           ``atom.coord`` float32, interactions.py:745
  hydrogen test         ``atom.element.strip() == 'H'``, interactions.py:712
  selection / water     ``atom in set(self.selection)``, ``get_full_id()[3][0] == 'W'``, :669-685
  radii                 ``atom.vdw_radius`` / ``atom.cov_radius`` (python floats), :717-718
  atom types            ``'hbond donor' in atom.atom_types`` ..., :779-921
  residue filters       ``res_bgn is res_end``; ``is_polypeptide``; ``hasattr(prev_residue/next_residue)``
                        and identity of those links, :726-741
  covalent test         ``ob.OBAtomAtomIter(ob_atom_bgn)`` ids, :750-754
  hydrogens             ``donor.h_coords`` float64, utils.py:86, :109, :145
  halogen neighbour     ``utils.get_single_bond_neighbour``, utils.py:612-635
  rings / amides        ``structure.rings[k]`` / ``structure.amides[k]`` dicts, :1071-1194, :1217-1382
"""
import importlib

import numpy as np

from . import abi
from .soa import AtomSoA, PlaneSoA

_TYPE_BITS = {k: 1 << i for i, k in enumerate(abi.ATOM_TYPE_KEYS)}


def _ob_module(ob=None):
    if ob is not None:
        return ob
    return importlib.import_module('openbabel.openbabel')


def single_bond_neighbour(ob, ob_atom):
    """utils.get_single_bond_neighbour (utils.py:612-635): first bonded atom over a single,
    non-aromatic bond that is not a hydrogen."""
    for bond in ob.OBAtomBondIter(ob_atom):
        if not (bond.GetBondOrder() == 1 and not bond.IsAromatic()):
            continue
        nbr = bond.GetNbrAtom(ob_atom)
        if nbr.GetAtomicNum() == 1:
            continue
        return nbr
    return None


class PackedComplex:
    """AtomSoA + plane SoAs of one complex, with the maps back to the host objects."""

    def __init__(self, atoms, soa, rings, ring_keys, amides, amide_keys, residues):
        self.atoms = atoms            # list of Bio atoms, index = SoA index
        self.soa = soa
        self.rings = rings            # PlaneSoA (float64)
        self.ring_keys = ring_keys    # SoA ring index -> key of structure.rings
        self.amides = amides          # PlaneSoA (float32)
        self.amide_keys = amide_keys
        self.residues = residues      # SoA residue index -> Bio residue


def pack_complex(ic, ob=None, inter_residue_bonds_only=True):
    """Pack ``ic`` (after ``initialize()`` and ``_make_selection()``) for the CUDA engine."""
    ob = _ob_module(ob)
    atoms = list(ic.selection_plus)
    n = len(atoms)
    index_of = {id(a): i for i, a in enumerate(atoms)}
    selection_set = set(ic.selection)

    xyz = np.empty((n, 3), dtype=np.float32)
    feat = np.zeros(n, dtype=np.uint32)
    res_id = np.empty(n, dtype=np.int32)
    rad_class = np.empty(n, dtype=np.uint16)
    rad_index = {}
    residues, res_index = [], {}
    h_counts = np.zeros(n, dtype=np.int32)
    h_list = []
    xnbr = None

    for i, a in enumerate(atoms):
        c = a.coord
        if c.dtype != np.float32:
            raise TypeError('atom.coord must be float32 (protein_reader.py:327)')
        xyz[i] = c
        f = 0
        for t in a.atom_types:
            f |= _TYPE_BITS.get(t, 0)
        if a.is_metal:
            f |= abi.F_IS_METAL
        if a.is_halogen:
            f |= abi.F_IS_HALOGEN
        if a.get_full_id()[3][0] == 'W':
            f |= abi.F_IS_WATER
        if a in selection_set:
            f |= abi.F_IN_SELECTION
        if a.element.strip() == 'H':
            f |= abi.F_ELEM_H
        if a.element == 'C':
            f |= abi.F_ELEM_C
        res = a.get_parent()
        if res.resname == 'MET' and a.element == 'S':
            f |= abi.F_MET_SULPHUR
        key = (float(a.vdw_radius), float(a.cov_radius))
        k = rad_index.get(key)
        if k is None:
            k = rad_index[key] = len(rad_index)
        rad_class[i] = k
        r = res_index.get(id(res))
        if r is None:
            r = res_index[id(res)] = len(residues)
            residues.append(res)
        res_id[i] = r
        if a.h_coords:
            h_counts[i] = len(a.h_coords)
            h_list.extend(a.h_coords)
        # the halogen / xbond-donor neighbour is only looked up by is_halogen_weak_hbond and is_xbond
        if (f & abi.F_XBOND_DONOR) or ((f & abi.F_IS_HALOGEN) and (f & abi.F_WEAK_HBOND_ACCEPTOR)):
            nb = single_bond_neighbour(ob, ic.ob_mol.GetAtomById(ic.bio_to_ob[a]))
            if nb is not None:
                if xnbr is None:
                    xnbr = np.zeros((n, 3), dtype=np.float32)
                xnbr[i] = ic.ob_to_bio[nb.GetId()].coord
                f |= abi.F_HAS_XNBR
        feat[i] = f

    rs = len(residues)
    res_prev = np.full(rs, -1, dtype=np.int32)
    res_next = np.full(rs, -1, dtype=np.int32)
    res_flags = np.zeros(rs, dtype=np.uint8)
    for r, res in enumerate(residues):
        fl = 0
        if getattr(res, 'is_polypeptide', False):
            fl |= abi.R_IS_POLYPEPTIDE
        if hasattr(res, 'prev_residue') and hasattr(res, 'next_residue'):
            fl |= abi.R_HAS_LINKS
            # a linked residue that has no atom in the list can never be `is` a listed one: -1
            if res.prev_residue is not None:
                res_prev[r] = res_index.get(id(res.prev_residue), -1)
            if res.next_residue is not None:
                res_next[r] = res_index.get(id(res.next_residue), -1)
        res_flags[r] = fl

    # covalent neighbours, restricted to listed atoms.  Pairs inside one residue never reach the
    # covalent test (interactions.py:729 precedes :750), so only inter-residue bonds can matter.
    b_off = np.zeros(n + 1, dtype=np.int32)
    b_nbr = []
    for i, a in enumerate(atoms):
        oba = ic.ob_mol.GetAtomById(ic.bio_to_ob[a])
        for nb in ob.OBAtomAtomIter(oba):
            other = ic.ob_to_bio.get(nb.GetId())
            j = index_of.get(id(other)) if other is not None else None
            if j is None:
                continue
            if inter_residue_bonds_only and res_id[j] == res_id[i]:
                continue
            b_nbr.append(j)
        b_off[i + 1] = len(b_nbr)

    vdw = np.empty(len(rad_index), dtype=np.float64)
    cov = np.empty(len(rad_index), dtype=np.float64)
    for (v, c), k in rad_index.items():
        vdw[k], cov[k] = v, c

    h_off = np.concatenate([[0], np.cumsum(h_counts)]).astype(np.int32)
    h_xyz = np.array(h_list, dtype=np.float64).reshape(-1, 3) if h_list else np.zeros((0, 3), dtype=np.float64)

    soa = AtomSoA(xyz=xyz, feat=feat, res_id=res_id, rad_class=rad_class, vdw=vdw, cov=cov,
                  res_prev=res_prev, res_next=res_next, res_flags=res_flags,
                  bond_off=b_off, bond_nbr=np.array(b_nbr, dtype=np.int32),
                  h_off=h_off, h_xyz=h_xyz, xnbr_xyz=xnbr)

    def plane_soa(groups, sel_ids, sel_plus_ids, is_f32):
        keys = list(groups)          # iteration order of the OrderedDict (:1071, :1218)
        m = len(keys)
        dt = np.float32 if is_f32 else np.float64
        center = np.zeros((m, 3), dtype=dt)
        normal = np.zeros((m, 3), dtype=dt)
        rid = np.full(m, -1, dtype=np.int32)
        flags = np.zeros(m, dtype=np.uint32)
        extra = {}
        for k, key in enumerate(keys):
            g = groups[key]
            if np.asarray(g['center']).dtype != dt or np.asarray(g['normal']).dtype != dt:
                raise TypeError('plane centre/normal dtype differs from the reference (%s)' % dt.__name__)
            center[k] = g['center']
            normal[k] = g['normal']
            res = g.get('residue')
            if res is not None:
                # `ring['residue'] == atom.get_parent()` (:981, :1091) is Entity.__eq__ (full id); residues
                # of one structure have distinct full ids, so equality coincides with identity of the
                # canonical object for that full id
                r = res_index.get(id(res))
                if r is None:
                    r = extra.get(res)
                    if r is None:
                        r = extra[res] = rs + len(extra)
                rid[k] = r
            if key in sel_ids:
                flags[k] |= abi.P_IN_SELECTION
            if key in sel_plus_ids:
                flags[k] |= abi.P_IN_SELECTION_PLUS
        return PlaneSoA(center, normal, rid, flags, is_f32), keys

    st = ic.biopython_str
    rings, ring_keys = plane_soa(getattr(st, 'rings', {}), set(ic.selection_ring_ids),
                                 set(ic.selection_plus_ring_ids), False)
    amides, amide_keys = plane_soa(getattr(st, 'amides', {}), set(ic.selection_amide_ids),
                                   set(ic.selection_plus_amide_ids), True)
    return PackedComplex(atoms, soa, rings, ring_keys, amides, amide_keys, residues)
