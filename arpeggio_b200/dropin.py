"""Drop-in for the contact engine of ``arpeggio.core.InteractionComplex``.

``CudaContactsMixin`` overrides the three protected methods through which
``InteractionComplex.run_arpeggio`` (arpeggio/core/interactions.py:329-347) reaches the contact
loops -- ``_calculate_atom_contacts`` (:693), ``_calculate_ring_contacts`` (:938) and
``_calculate_group_contacts`` (:1208) -- and fills the same result bags (``atom_contacts``,
``plane_plane_contacts``, ``atom_plane_contacts``, ``group_group_contacts``,
``group_plane_contacts``, :84-88) with the reference's own namedtuples, so that
``get_contacts()`` (:172-212) and everything downstream run unchanged.  Everything before the
loops (file parsing, typing, ``_make_selection``) stays the reference's host code.

    from arpeggio_b200.dropin import cuda_interaction_complex
    InteractionComplex = cuda_interaction_complex()        # subclass of arpeggio.core.InteractionComplex
    ic = InteractionComplex('1tqn_h.cif'); ic.structure_checks(); ic.initialize()
    ic.run_arpeggio(['/A/508/'], 5.0, 0.1, False); contacts = ic.get_contacts()

Side effects of the loops: the per-atom SIFt words, integer SIFts and hbond / polar counters
(interactions.py:822-852, :924-934) come from the GPU reduction ``ContactEngine.atom_sifts`` and are
written back onto the atoms as the same attributes (``atom.sift``, ``integer_sift_inter_only``,
``actual_fsift``, ``actual_hbonds`` ...); the per-residue counters of the plane loops (:1040-1057,
:1171-1176, :1290-1291, :1371-1373) are re-derived from the plane records.  Every run starts them
from zero (what ``initialize()`` leaves), so ``write_atom_sifts`` / ``_calc_residue_sifts`` and the other
CSV writers work on top.  List order: the reference emits pairs in KD-tree traversal order; here
records come out sorted by (bgn index, end index) of ``selection_plus`` -- that is also the loop order
the order-dependent ``integer_sift`` (utils.py:233) is evaluated for.
"""
import collections
import collections.abc

import numpy as np

from . import abi, jsonout, params as arp_params
from .engine import ContactEngine
from .packing import pack_complex

# the reference's record types (interactions.py:19-29); the real ones are used when importable
AtomPlaneContact = collections.namedtuple('AtomPlaneContact',
                                          ['bgn_atom', 'end_res', 'end_res_atoms', 'distance', 'sifts', 'text'])
PlanePlaneContact = collections.namedtuple('PlanePlaneContact',
                                           ['bgn_id', 'bgn_res', 'bgn_res_atoms', 'end_id', 'end_res', 'end_res_atoms',
                                            'distance', 'contact_type', 'text'])
AtomAtomContact = collections.namedtuple('AtomAtomContact', ['bgn_atom', 'end_atom', 'sifts', 'contact_type', 'distance'])

_ENGINES = {}


def shared_engine(device=0):
    """One ContactEngine per device and process (a context is not re-entrant: callers that drive
    several complexes from several threads should give every thread its own engine instead)."""
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = ContactEngine(device)
    return eng


def _record_types(obj):
    """The namedtuple classes of the module the host class comes from (so that isinstance / pickling
    by reference users keep working), else the local mirrors."""
    import sys
    for klass in type(obj).__mro__:
        mod = sys.modules.get(klass.__module__)
        if mod is not None and all(hasattr(mod, n) for n in ('AtomAtomContact', 'PlanePlaneContact', 'AtomPlaneContact')):
            return mod.AtomAtomContact, mod.PlanePlaneContact, mod.AtomPlaneContact
    return AtomAtomContact, PlanePlaneContact, AtomPlaneContact


def _contact_types_of(obj):
    """config.CONTACT_TYPES of the reference the host class belongs to (single source of truth for the
    thresholds, config.py:592-660), else the defaults mirrored in params.py."""
    import sys
    for klass in type(obj).__mro__:
        mod = sys.modules.get(klass.__module__)
        cfg = getattr(mod, 'config', None) if mod is not None else None
        if cfg is not None and hasattr(cfg, 'CONTACT_TYPES'):
            return cfg.CONTACT_TYPES, getattr(cfg, 'CONTACT_TYPES_DIST_MAX', arp_params.DEFAULT_DIST_MAX), \
                getattr(cfg, 'VDW_RADII', {}).get('H', arp_params.DEFAULT_H_VDW)
    return None, arp_params.DEFAULT_DIST_MAX, arp_params.DEFAULT_H_VDW


class LazyAtomContacts(collections.abc.Sequence):
    """``atom_contacts`` without one namedtuple per record up front: a read-only sequence over the
    record array that builds the reference's ``AtomAtomContact(bgn_atom, end_atom, sifts, contact_type,
    distance)`` (interactions.py:28-29, :936) when an element is asked for.  Iteration, indexing, slicing,
    ``len`` and ``filter(...)`` -- everything the reference does with the list after the loop -- work."""

    def __init__(self, records, atoms, record_type):
        self.records, self._atoms, self._type = records, atoms, record_type

    def __len__(self):
        return self.records.shape[0]

    def _make(self, r):
        m = int(r['mask'])
        return self._type(self._atoms[int(r['i'])], self._atoms[int(r['j'])], [m >> b & 1 for b in range(abi.SIFT_NBITS)],
                          abi.CLASS_NAMES[(m >> abi.CLASS_SHIFT) & abi.CLASS_MASK], r['dist'])

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._make(r) for r in self.records[k]]
        return self._make(self.records[k])

    def __iter__(self):
        atoms, make = self._atoms, self._type
        rec = self.records
        names = abi.CLASS_NAMES
        dist = rec['dist']
        for k, (i, j, m) in enumerate(zip(rec['i'].tolist(), rec['j'].tolist(), rec['mask'].tolist())):
            yield make(atoms[i], atoms[j], [m >> b & 1 for b in range(abi.SIFT_NBITS)],
                       names[(m >> abi.CLASS_SHIFT) & abi.CLASS_MASK], dist[k])


def _host_helpers(obj):
    """utils.make_pymol_json / utils.get_residue_name of the module the host class comes from (utils.py:530-564,
    :748-767)."""
    import sys
    for klass in type(obj).__mro__:
        mod = sys.modules.get(klass.__module__)
        if mod is None:
            continue
        for holder in (getattr(mod, 'utils', None), mod):
            if holder is not None and hasattr(holder, 'make_pymol_json') and hasattr(holder, 'get_residue_name'):
                return holder.make_pymol_json, holder.get_residue_name
    raise AttributeError('the host class offers no make_pymol_json / get_residue_name')


MAINCHAIN_ATOMS = frozenset(('N', 'C', 'CA', 'O', 'OXT'))          # config.py:35

_RESIDUE_PLANE_SIFTS = (('ring_ring_inter_integer_sift', 9), ('ring_atom_inter_integer_sift', 5),
                        ('atom_ring_inter_integer_sift', 5), ('mc_atom_ring_inter_integer_sift', 5),
                        ('sc_atom_ring_inter_integer_sift', 5), ('amide_ring_inter_integer_sift', 1),
                        ('ring_amide_inter_integer_sift', 1), ('amide_amide_inter_integer_sift', 1))


def _bump(residue, name, k):
    """residue.<name>[k] += 1 as the reference spells it (a fresh list each time, interactions.py:1047)."""
    v = list(getattr(residue, name))
    v[k] = v[k] + 1
    setattr(residue, name, v)


def apply_atom_sifts(atoms, sifts, integer_sifts=False):
    """Write an ``arp_atom_sift`` array back onto the atoms as the attributes the reference's pair loop
    leaves (utils.py:182-242: sift*, integer_sift*, actual_fsift*; interactions.py:822-852: actual_hbonds*,
    actual_polars*), as plain Python lists / ints.

    ``integer_sift*`` is ORDER-DEPENDENT in the reference: utils.py:233 ASSIGNS sift-before-this-contact + SIFt at
    every contact, so the value is that of the atom's last contact in loop order, and the reference's loop order is
    the KD-tree traversal of Bio.PDB.NeighborSearch.search_all, which this library does not reproduce (records are
    (bgn, end)-sorted).  integer_sifts=False (the default of the mixin) therefore REMOVES the attributes, so that a
    consumer (_calc_residue_sifts, the SIFt CSV writers) fails loudly instead of reading values that can differ from
    a BioPython run; integer_sifts=True writes the values of the sorted loop order."""
    nb = abi.SIFT_NBITS
    shifts = np.arange(nb, dtype=np.uint32)
    for c, suffix in enumerate(abi.SIFT_CATEGORIES):
        bits = ((sifts['sift'][:, c].astype(np.uint32)[:, None] >> shifts) & 1).astype(np.int64).tolist()
        integer = ((sifts['integer_sift'][:, c][:, None] >> (2 * shifts)) & 3).astype(np.int64).tolist()
        hb, pl = sifts['hbonds'][:, c].tolist(), sifts['polars'][:, c].tolist()
        for k, atom in enumerate(atoms):
            setattr(atom, 'sift' + suffix, bits[k])
            if integer_sifts:
                setattr(atom, 'integer_sift' + suffix, integer[k])
            elif hasattr(atom, 'integer_sift' + suffix):
                delattr(atom, 'integer_sift' + suffix)
            setattr(atom, 'actual_fsift' + suffix, bits[k][5:])
            setattr(atom, 'actual_hbonds' + suffix, hb[k])
            setattr(atom, 'actual_polars' + suffix, pl[k])


class CudaContactsMixin:
    """Mix in before ``arpeggio.core.InteractionComplex`` (or a duck-typed stand-in)."""

    cuda_device = 0
    cuda_engine = None          # set to a private ContactEngine to avoid the shared one
    cuda_atom_sifts = True      # reproduce the per-atom / per-residue SIFt side effects of the loops
    cuda_make_selection = True  # False: keep the host class's own _make_selection (the reference's list(set) order of
                                # selection_plus depends on its KD-tree pair order, INTEGRATION.md "differences" 3)
    cuda_integer_sifts = False  # opt in to atom.integer_sift* evaluated in (bgn, end)-sorted loop order: the reference's
                                # value depends on Bio.PDB's KD-tree traversal order (utils.py:233), see apply_atom_sifts
    cuda_lazy_contacts = False  # atom_contacts as a LazyAtomContacts sequence instead of a list of namedtuples

    def _cuda_reset_residue_sifts(self, names):
        """The counters start from what _initialize_residue_sift leaves (interactions.py:1869-1883)."""
        sizes = dict(_RESIDUE_PLANE_SIFTS)
        for residue in self.biopython_str.get_residues():
            for n in names:
                setattr(residue, n, [0] * sizes[n])

    # ------------------------------------------------------------------
    def _cuda_engine(self):
        return self.cuda_engine if self.cuda_engine is not None else shared_engine(self.cuda_device)

    def _cuda_packed(self):
        """SoA image of the current selection (rebuilt whenever _make_selection produced new lists; call
        cuda_invalidate() after changing coordinates, hydrogens, atom types, rings or amides in place)."""
        key = (id(self.selection_plus), len(self.selection_plus), id(self.selection), len(self.selection),
               getattr(self, '_cuda_pack_version', 0))
        cached = getattr(self, '_cuda_pack_cache', None)
        if cached is None or cached[0] != key:
            cached = (key, pack_complex(self, ob=getattr(self, '_cuda_ob_module', None)))
            self._cuda_pack_cache = cached
        return cached[1]

    def cuda_invalidate(self):
        """Forget the packed image: the next contact call packs the complex again (coordinates after
        minimize_hydrogens, atom types, rings, amides ... are read at packing time)."""
        self._cuda_pack_version = getattr(self, '_cuda_pack_version', 0) + 1
        self._cuda_pack_cache = None

    def _cuda_params(self, interacting_cutoff=None, vdw_comp_factor=None, include_sequence_adjacent=None):
        last = getattr(self, '_cuda_last_run', None)
        if interacting_cutoff is None:
            if last is None:
                p = getattr(self, 'params', None)      # Parameters namedtuple of __init__ (interactions.py:41-43)
                last = (getattr(p, 'interacting_threshold', 5.0), getattr(p, 'vdw_comp_factor', 0.1), False)
            interacting_cutoff, vdw_comp_factor, include_sequence_adjacent = last
        ct, dist_max, h_vdw = _contact_types_of(self)
        return arp_params.make_params(interacting_cutoff, vdw_comp_factor, include_sequence_adjacent,
                                      contact_types=ct, dist_max=dist_max, h_vdw=h_vdw)

    # ------------------------------------------------------------------
    def _cuda_parse_selection(self, selections, entity):
        """utils.selection_parser of the host module (interactions.py:1396)."""
        import sys
        for klass in type(self).__mro__:
            mod = sys.modules.get(klass.__module__)
            utils = getattr(mod, 'utils', None) if mod is not None else None
            if utils is not None and hasattr(utils, 'selection_parser'):
                return utils.selection_parser(selections, entity)
        raise AttributeError('the host class offers no utils.selection_parser')

    def _make_selection(self, selections):
        """Replaces interactions.py:1384-1451: the binding-site expansion -- every atom within 6 A of
        a selected atom -- is one GPU pass (``arp_flag_within``) instead of ``search_all(6.0)`` over the whole
        structure with a Python tuple per pair.  The bookkeeping (residue / ring / amide id sets) follows the
        reference.  ``selection_plus`` keeps the reference's construction, ``list(set(...))``: like there, its
        order is whatever the set yields."""
        import logging
        import sys
        from .soa import AtomSoA
        if not self.cuda_make_selection:
            self.cuda_invalidate()
            return super()._make_selection(selections)
        entity = list(self.s_atoms)
        selection = entity if not selections else self._cuda_parse_selection(selections, entity)
        if not selection:
            logging.error('Selection was empty.')
            raise AttributeError('Selection must not be empty.')
        rings, amides = self.biopython_str.rings, self.biopython_str.amides
        selection_set = set(selection)
        n = len(entity)
        feat = np.zeros(n, dtype=np.uint32)
        feat[[k for k, a in enumerate(entity) if a in selection_set]] = abi.F_IN_SELECTION
        soa = AtomSoA(xyz=np.array([a.coord for a in entity], dtype=np.float32).reshape(-1, 3), feat=feat,
                      res_id=np.zeros(n, np.int32), rad_class=np.zeros(n, np.uint16), vdw=np.ones(1), cov=np.ones(1),
                      res_prev=np.full(1, -1, np.int32), res_next=np.full(1, -1, np.int32), res_flags=np.zeros(1, np.uint8))
        eng = self._cuda_engine()
        eng.upload_atoms(soa)
        near = eng.flag_within(6.0)
        selection_plus = set(selection)
        selection_plus.update(a for a, f in zip(entity, near.tolist()) if f)
        selection_plus = list(selection_plus)

        selection_residues = {a.get_parent() for a in selection}
        selection_plus_residues = {a.get_parent() for a in selection_plus}
        self.cuda_invalidate()          # new lists; the engine also holds the whole structure now, not the packed selection
        self.selection = selection
        self.selection_ring_ids = {k for k in rings if rings[k]['residue'] in selection_residues}
        self.selection_amide_ids = {k for k in amides if amides[k]['residue'] in selection_residues}
        self.selection_plus = selection_plus
        self.selection_plus_residues = selection_plus_residues
        self.selection_plus_ring_ids = {k for k in rings if rings[k]['residue'] in selection_plus_residues}
        self.selection_plus_amide_ids = {k for k in amides if amides[k]['residue'] in selection_plus_residues}
        for klass in type(self).__mro__:                 # the reference leaves a search tree over selection_plus
            mod = sys.modules.get(klass.__module__)
            ns = getattr(mod, 'NeighborSearch', None) if mod is not None else None
            if ns is not None:
                self.ns = ns(selection_plus)
                break

    def _assign_aromatic_rings_to_residues(self):
        """Replaces interactions.py:1453-1492: the closest atom within 3 A of every ring centroid
        comes from the GPU (ties: lowest atom index); the bookkeeping on rings and residues is the reference's."""
        import logging
        import sys
        for klass in type(self).__mro__:                 # the reference also leaves a whole-structure search tree
            mod = sys.modules.get(klass.__module__)
            ns = getattr(mod, 'NeighborSearch', None) if mod is not None else None
            if ns is not None:
                self.ns = ns(self.s_atoms)
                break
        rings = self.biopython_str.rings
        keys = list(rings)
        if not keys:
            return
        atoms = list(self.s_atoms)
        xyz = np.array([a.coord for a in atoms], dtype=np.float32).reshape(-1, 3)
        centers = np.array([rings[k]['center'] for k in keys], dtype=np.float64).reshape(-1, 3)
        eng = self._cuda_engine()
        nearest, dist = eng.ring_nearest_atom(xyz, centers, 3.0)
        for k, a, d in zip(keys, nearest.tolist(), dist):
            if a < 0:
                logging.warning(f'Residue assignment was not possible for ring {k}.')
                rings[k]['residue'] = None
                continue
            residue = atoms[a].get_parent()
            rings[k]['residue'] = residue
            rings[k]['residue_shortest_distance'] = d
            if not hasattr(residue, 'rings'):
                residue.rings = []
            residue.rings.append(k)

    def _calculate_atom_contacts(self, interacting_cutoff, vdw_comp_factor, include_sequence_adjacent):
        """Replaces interactions.py:693-936 (neighbour search + per-pair rules) with the CUDA path."""
        AAC, _, _ = _record_types(self)
        self._cuda_last_run = (interacting_cutoff, vdw_comp_factor, include_sequence_adjacent)
        packed = self._cuda_packed()
        eng = self._cuda_engine()
        eng.set_params(self._cuda_params(interacting_cutoff, vdw_comp_factor, include_sequence_adjacent))
        rec = eng.pairs(packed.soa, sorted=True)
        if rec.shape[0] and np.any(rec['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR)):
            # utils.is_xbond dereferences None when the donor has no single-bond neighbour (utils.py:173)
            raise AttributeError("'NoneType' object has no attribute 'coord'")
        atoms = packed.atoms
        self._cuda_pair_records = rec
        lazy = LazyAtomContacts(rec, atoms, AAC)
        self.atom_contacts = lazy if self.cuda_lazy_contacts else list(lazy)
        if self.cuda_atom_sifts:
            apply_atom_sifts(atoms, eng.atom_sifts(), integer_sifts=self.cuda_integer_sifts)

    def _calculate_ring_contacts(self):
        """Replaces interactions.py:938-1194 (plane-plane and atom-plane)."""
        _, PPC, APC = _record_types(self)
        packed = self._cuda_packed()
        eng = self._cuda_engine()
        eng.set_params(self._cuda_params())
        eng.upload_atoms(packed.soa)
        eng.upload_planes(packed.rings, packed.amides)
        # all four plane terms in one launch sequence; the amide terms are kept for _calculate_group_contacts
        terms = eng.planes_all()
        self._cuda_plane_terms = (id(packed), terms)
        rings = self.biopython_str.rings
        names = {}

        def ring_atoms(key):
            v = names.get(key)
            if v is None:
                v = names[key] = sorted(a.get_id() for a in rings[key]['atoms'])
            return v

        sifts = self.cuda_atom_sifts
        if sifts:
            self._cuda_reset_residue_sifts(('ring_ring_inter_integer_sift', 'ring_atom_inter_integer_sift',
                                            'atom_ring_inter_integer_sift', 'mc_atom_ring_inter_integer_sift',
                                            'sc_atom_ring_inter_integer_sift'))
        inter = abi.CLASS_NAMES.index('INTER')
        self.plane_plane_contacts = []
        for r in terms['ring_ring']:
            ka, kb = packed.ring_keys[int(r['a'])], packed.ring_keys[int(r['b'])]
            code = int(r['code'])
            labels = [abi.GEOM_NAMES[code & 0xF]]
            if (code >> 4) & 0xF != 0xF:
                labels.append(abi.GEOM_NAMES[(code >> 4) & 0xF])
            if sifts and (code >> 8) & 7 == inter and not code >> 11 & 1:
                # interactions.py:1171-1176, once per VISIT: (a, b) credits ring a's residue with its geometry,
                # (b, a) credits ring b's with its own -- the second label when it differed, else the same
                g1 = code & 0xF
                g2 = (code >> 4) & 0xF
                if g1 < 9:
                    _bump(rings[ka]['residue'], 'ring_ring_inter_integer_sift', g1)
                g2 = g1 if g2 == 0xF else g2
                if g2 < 9:
                    _bump(rings[kb]['residue'], 'ring_ring_inter_integer_sift', g2)
            self.plane_plane_contacts.append(PPC(ka, rings[ka]['residue'], list(ring_atoms(ka)), kb, rings[kb]['residue'],
                                                 list(ring_atoms(kb)), np.float64(r['dist']), labels,
                                                 abi.CLASS_NAMES[(code >> 8) & 7]))
        self.atom_plane_contacts = []
        for r in terms['atom_ring']:
            key = packed.ring_keys[int(r['ring'])]
            code = int(r['code'])
            labels = sorted(n for b, n in enumerate(abi.AP_NAMES) if code >> b & 1)
            if sifts and (code >> 8) & 7 == inter and not code >> 11 & 1:          # interactions.py:1039-1057
                atom = packed.atoms[int(r['atom'])]
                parent = atom.get_parent()
                for k in range(len(abi.AP_NAMES)):
                    if code >> k & 1:
                        _bump(rings[key]['residue'], 'ring_atom_inter_integer_sift', k)
                        _bump(parent, 'atom_ring_inter_integer_sift', k)
                        if parent in self.polypeptide_residues:
                            _bump(parent, ('mc' if atom.name in MAINCHAIN_ATOMS else 'sc') + '_atom_ring_inter_integer_sift', k)
            self.atom_plane_contacts.append(APC(packed.atoms[int(r['atom'])], rings[key]['residue'], list(ring_atoms(key)),
                                                np.float64(r['dist']), labels, abi.CLASS_NAMES[(code >> 8) & 7]))

    def _calculate_group_contacts(self):
        """Replaces interactions.py:1208-1382 (amide-amide and amide-ring)."""
        _, PPC, _ = _record_types(self)
        packed = self._cuda_packed()
        eng = self._cuda_engine()
        stash = getattr(self, '_cuda_plane_terms', None)
        self._cuda_plane_terms = None
        if stash is not None and stash[0] == id(packed):
            terms = stash[1]                                  # computed together with the ring terms just before
        else:
            eng.set_params(self._cuda_params())
            eng.upload_planes(packed.rings, packed.amides)
            terms = {'amide_amide': eng.amide_amide(), 'amide_ring': eng.amide_ring()}
        rings, amides = self.biopython_str.rings, self.biopython_str.amides

        def names(group):
            return sorted(a.get_id() for a in group['atoms'])

        sifts = self.cuda_atom_sifts
        if sifts:
            self._cuda_reset_residue_sifts(('amide_ring_inter_integer_sift', 'ring_amide_inter_integer_sift',
                                            'amide_amide_inter_integer_sift'))
        inter = abi.CLASS_NAMES.index('INTER')

        def counted(code):                  # contact_type == 'INTER' and not intra_residue (interactions.py:1290, :1371)
            return sifts and (code >> 8) & 7 == inter and not code >> 11 & 1

        self.group_group_contacts = []
        for r in terms['amide_amide']:
            a, b = amides[packed.amide_keys[int(r['a'])]], amides[packed.amide_keys[int(r['b'])]]
            if counted(int(r['code'])):
                _bump(a['residue'], 'amide_amide_inter_integer_sift', 0)
            self.group_group_contacts.append(PPC(a['amide_id'], a['residue'], names(a), b['amide_id'], b['residue'], names(b),
                                                 np.float32(r['dist']), ['AMIDEAMIDE'],
                                                 abi.CLASS_NAMES[(int(r['code']) >> 8) & 7]))
        self.group_plane_contacts = []
        for r in terms['amide_ring']:
            a, g = amides[packed.amide_keys[int(r['a'])]], rings[packed.ring_keys[int(r['b'])]]
            if counted(int(r['code'])):
                _bump(a['residue'], 'amide_ring_inter_integer_sift', 0)
                _bump(g['residue'], 'ring_amide_inter_integer_sift', 0)
            self.group_plane_contacts.append(PPC(a['amide_id'], a['residue'], names(a), g['ring_id'], g['residue'], names(g),
                                                 np.float64(r['dist']), ['AMIDERING'],
                                                 abi.CLASS_NAMES[(int(r['code']) >> 8) & 7]))


    # ------------------------------------------------------------------
    def _cuda_json_parts(self, threads=None):
        rec = getattr(self, '_cuda_pair_records', None)
        if rec is None or len(self.atom_contacts) != rec.shape[0]:
            raise RuntimeError('the contact JSON needs the record stream of the last _calculate_atom_contacts')
        make_json, residue_name = _host_helpers(self)
        frags = []
        for atom in self._cuda_packed().atoms:
            d = make_json(atom)
            d['label_comp_type'] = self.component_types[residue_name(atom)]
            frags.append(jsonout.atom_fragment(d))
        body = jsonout.pairs_json(rec, frags, threads=threads)
        keep = self.atom_contacts
        self.atom_contacts = []
        try:
            others = self.get_contacts()
        finally:
            self.atom_contacts = keep
        return body, others

    def contacts_json_text(self, threads=None):
        """The text ``json.dump(self.get_contacts(), fp, indent=4, sort_keys=True)`` writes
        (process_protein_cli.py:187-188), byte for byte, without a Python dict per atom-atom contact: the records
        of the last run go through the C emitter, the plane / group entries through the host's own get_contacts."""
        body, others = self._cuda_json_parts(threads)
        return jsonout.splice(body, others)

    def write_contacts_json(self, path, threads=None):
        body, others = self._cuda_json_parts(threads)
        with open(path, 'wb') as fp:
            jsonout.write_spliced(fp, body, others)


def cuda_interaction_complex(base_cls=None, device=0):
    """A subclass of ``arpeggio.core.InteractionComplex`` (or of ``base_cls``) whose contact engine is
    libarpeggio_cuda.so.  Importing the reference needs BioPython, OpenBabel and gemmi."""
    if base_cls is None:
        from arpeggio.core import InteractionComplex as base_cls   # noqa: N813
    return type('CudaInteractionComplex', (CudaContactsMixin, base_cls), {'cuda_device': device})
