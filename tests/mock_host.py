"""A stand-in host class for the drop-in tests on boxes without the reference (GPU box): it holds what
``InteractionComplex`` holds after ``initialize()`` + ``_make_selection()`` and mirrors
``run_arpeggio`` (interactions.py:329-347) and ``get_contacts`` (:172-212, :2063-2113; utils.py:530-564,
:748-767) verbatim, so that the CUDA mixin can be driven exactly as the reference drives its own
methods and its JSON compared with the JSON the real reference produced (tests/golden)."""
from functools import reduce

import numpy as np

import mockbio
from arpeggio_b200.dropin import CudaContactsMixin


def make_pymol_json(entity):                      # utils.py:530-564
    if isinstance(entity, mockbio.Atom):
        residue = entity.get_parent()
        chain = residue.get_parent()
        return {'label_comp_id': residue.resname, 'auth_seq_id': residue.id[1], 'auth_asym_id': chain.id,
                'auth_atom_id': entity.name, 'pdbx_PDB_ins_code': residue.id[2]}
    if isinstance(entity, mockbio.Residue):
        chain = entity.get_parent()
        return {'label_comp_id': entity.resname, 'auth_seq_id': entity.id[1], 'auth_asym_id': chain.id,
                'pdbx_PDB_ins_code': entity.id[2]}
    raise TypeError('Cannot make a json object from non-Atom/Residue object.')


def get_residue_name(entity):                     # utils.py:748-767
    return entity.get_parent().get_resname() if isinstance(entity, mockbio.Atom) else entity.get_resname()


class MockHost(CudaContactsMixin, mockbio.MockComplex):
    """mockbio.MockComplex + the reference's run/export surface + the CUDA contact engine."""

    _cuda_ob_module = mockbio.make_ob_module()

    def __init__(self, cx):
        self.__dict__.update(cx.__dict__)

    def run_arpeggio(self, interacting_cutoff, vdw_comp_factor, include_sequence_adjacent):   # interactions.py:342-347
        self._calculate_atom_contacts(interacting_cutoff, vdw_comp_factor, include_sequence_adjacent)
        self._calculate_ring_contacts()
        self._calculate_group_contacts()

    def get_contacts(self):                       # interactions.py:172-212
        contacts = ['clash', 'covalent', 'vdw_clash', 'vdw', 'proximal', 'hbond', 'weak_hbond', 'xbond', 'ionic',
                    'metal_complex', 'aromatic', 'hydrophobic', 'carbonyl', 'polar', 'weak_polar']
        bag = []
        for c in self.atom_contacts:
            e = {'bgn': make_pymol_json(c.bgn_atom), 'end': make_pymol_json(c.end_atom)}
            e['bgn']['label_comp_type'] = self.component_types[get_residue_name(c.bgn_atom)]
            e['end']['label_comp_type'] = self.component_types[get_residue_name(c.end_atom)]
            e['type'] = 'atom-atom'
            e['distance'] = round(np.float64(c.distance), 2)
            e['contact'] = [k for k, v in zip(contacts, c.sifts) if v == 1]
            e['interacting_entities'] = c.contact_type
            bag.append(e)
        for c in self.plane_plane_contacts:
            bag.append(self._plane_plane(c, 'plane-plane'))
        for c in self.atom_plane_contacts:
            bag.append(self._atom_plane(c, 'atom-plane'))
        for c in self.group_group_contacts:
            bag.append(self._plane_plane(c, 'group-group'))
        for c in self.group_plane_contacts:
            bag.append(self._plane_plane(c, 'group-plane'))
        return bag

    def _plane_plane(self, c, kind):              # interactions.py:2063-2087
        e = {'bgn': make_pymol_json(c.bgn_res), 'end': make_pymol_json(c.end_res)}
        e['bgn']['label_comp_type'] = self.component_types[get_residue_name(c.bgn_res)]
        e['bgn']['auth_atom_id'] = reduce(lambda l, m: f'{l},{m}', c.bgn_res_atoms)
        e['end']['label_comp_type'] = self.component_types[get_residue_name(c.end_res)]
        e['end']['auth_atom_id'] = reduce(lambda l, m: f'{l},{m}', c.end_res_atoms)
        e['type'] = kind
        e['distance'] = round(np.float64(c.distance), 2)
        e['contact'] = c.contact_type
        e['interacting_entities'] = c.text
        return e

    def _atom_plane(self, c, kind):               # interactions.py:2090-2113
        e = {'bgn': make_pymol_json(c.bgn_atom), 'end': make_pymol_json(c.end_res)}
        e['bgn']['label_comp_type'] = self.component_types[get_residue_name(c.bgn_atom)]
        e['end']['auth_atom_id'] = reduce(lambda l, m: f'{l},{m}', c.end_res_atoms)
        e['end']['label_comp_type'] = self.component_types[get_residue_name(c.end_res)]
        e['type'] = kind
        e['distance'] = round(np.float64(c.distance), 2)
        e['contact'] = c.sifts
        e['interacting_entities'] = c.text
        return e


def host_from_golden(g):
    """Rebuild the mock complex of a golden fixture from its recipe and restore the selection lists
    exactly as the reference's _make_selection left them (order matters: index = list position)."""
    cx = mockbio.build_complex(**g.meta['recipe'])
    by_serial = {a.serial_number: a for a in cx.s_atoms}
    host = MockHost(cx)
    host.selection = [by_serial[s] for s in g.meta['selection_serials']]
    host.selection_plus = [by_serial[s] for s in g.meta['selection_plus_serials']]
    host.selection_ring_ids = list(g.meta['selection_ring_ids'])
    host.selection_plus_ring_ids = list(g.meta['selection_plus_ring_ids'])
    host.selection_amide_ids = list(g.meta['selection_amide_ids'])
    host.selection_plus_amide_ids = list(g.meta['selection_plus_amide_ids'])
    return host
