"""Duck-typed stand-ins for the BioPython / OpenBabel objects the contact engine reads.

BioPython, OpenBabel and gemmi are not installable in this environment, so the
tests build "mini complexes" out of these classes.  They implement exactly the
attribute surface the reference touches on the hot path
(arpeggio/core/interactions.py:643-1451, utils.py:73-179, :530-564, :612-635):

    Atom      .coord (float32[3]) .element .name .serial_number .get_parent() .get_full_id()
              .get_id() plus the attributes initialize() decorates atoms with
              (.atom_types .h_coords .vdw_radius .cov_radius .is_metal .is_halogen)
    Residue   .resname .id .child_list .get_parent() .get_resname() .is_polypeptide
              [.prev_residue .next_residue] (only polypeptide residues, interactions.py:1687-1693)
    Chain     .id
    OBAtom    .GetId() .GetAtomicNum(); OBBond .GetBondOrder() .IsAromatic() .GetNbrAtom()
    OBMol     .GetAtomById()
    ob module OBAtomAtomIter, OBAtomBondIter
    NeighborSearch(atom_list) .search_all(r) .search(center, r)   (Bio.PDB.NeighborSearch,
              restated: double coordinates, d2 <= r*r, index1 < index2; emission order
              here is ascending (i, j), the KD-tree order of the real one is not reproducible)

``install_stubs()`` registers these under the module names the reference imports
so that ``tests/golden/make_golden.py`` can import and run the *real* reference
methods on a mock complex.  ``build_complex()`` is the seeded generator; it does
not depend on the reference.
"""
import collections
import sys
import types

import numpy as np


# --------------------------------------------------------------------------
# Bio.PDB look-alikes
# --------------------------------------------------------------------------
class Chain:
    def __init__(self, cid):
        self.id = cid
        self.child_list = []

    def get_parent(self):
        return None


class Residue:
    def __init__(self, chain, hetflag, resseq, icode, resname):
        self.parent = chain
        self.id = (hetflag, resseq, icode)
        self.resname = resname
        self.child_list = []
        self.is_polypeptide = False          # interactions.py:1860
        chain.child_list.append(self)

    def get_parent(self):
        return self.parent

    def get_resname(self):
        return self.resname

    def get_full_id(self):
        return ('structure', 0, self.parent.id, self.id)

    # Bio.PDB.Entity compares and hashes by full id
    def __eq__(self, other):
        return isinstance(other, Residue) and self.get_full_id() == other.get_full_id()

    def __hash__(self):
        return hash(self.get_full_id())

    def __repr__(self):
        return f'<Residue {self.resname} het={self.id[0]} resseq={self.id[1]} icode={self.id[2]}>'


class Atom:
    def __init__(self, residue, name, element, coord, serial):
        self.parent = residue
        self.name = name
        self.element = element
        self.coord = np.array(coord, dtype='f')   # protein_reader.py:327
        self.serial_number = serial
        self.altloc = ' '
        residue.child_list.append(self)
        # what InteractionComplex.initialize() adds
        self.atom_types = set()
        self.h_coords = []
        self.vdw_radius = 0.0
        self.cov_radius = 0.0
        self.is_metal = False
        self.is_halogen = False

    def get_parent(self):
        return self.parent

    def get_id(self):
        return self.name

    def get_coord(self):
        return self.coord

    def get_full_id(self):
        r = self.parent
        return ('structure', 0, r.parent.id, r.id, (self.name, self.altloc))

    def __eq__(self, other):
        return isinstance(other, Atom) and self.get_full_id()[1:] == other.get_full_id()[1:]

    def __hash__(self):
        return hash(self.get_full_id())

    def __repr__(self):
        return f'<Atom {self.name}>'


class DisorderedAtom:   # only needed for isinstance() in InteractionComplex.__init__
    pass


class NeighborSearch:
    """Bio.PDB.NeighborSearch restated (see module docstring).

    pair_order_seed: Bio.PDB reports the pairs of search_all in KD-tree traversal order, which nothing here can
    reproduce; a seed makes search_all return its pairs in a seeded shuffle instead of ascending (i, j), to exercise
    everything that depends on that order (the insertion order of the reference's selection_plus set, the
    order-dependent integer_sift, the order of atom_contacts)."""

    pair_order_seed = None

    def __init__(self, atom_list, bucket_size=10):
        self.atom_list = list(atom_list)
        self.coords = np.array([a.get_coord() for a in self.atom_list], dtype='d').reshape(-1, 3)

    def search_all(self, radius, level='A'):
        c = self.coords
        r2 = radius * radius
        out = []
        for i in range(len(c) - 1):
            d = c[i + 1:] - c[i]
            s = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
            for j in np.nonzero(s <= r2)[0]:
                out.append((self.atom_list[i], self.atom_list[i + 1 + int(j)]))
        if self.pair_order_seed is not None:
            order = np.random.default_rng(self.pair_order_seed).permutation(len(out))
            out = [out[int(k)] for k in order]
        return out

    def search(self, center, radius, level='A'):
        center = np.require(center, dtype='d')
        d = self.coords - center
        s = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        return [self.atom_list[int(i)] for i in np.nonzero(s <= radius * radius)[0]]


# --------------------------------------------------------------------------
# OpenBabel look-alikes
# --------------------------------------------------------------------------
class OBBond:
    def __init__(self, a, b, order=1, aromatic=False):
        self.a, self.b, self.order, self.aromatic = a, b, order, aromatic

    def GetBondOrder(self):
        return self.order

    def IsAromatic(self):
        return self.aromatic

    def GetNbrAtom(self, atom):
        return self.b if atom is self.a else self.a


class OBAtom:
    def __init__(self, oid, atomic_num):
        self.oid, self.atomic_num = oid, atomic_num
        self.bonds = []

    def GetId(self):
        return self.oid

    def GetAtomicNum(self):
        return self.atomic_num


class OBMol:
    def __init__(self):
        self.atoms = {}

    def add_atom(self, oid, atomic_num):
        a = OBAtom(oid, atomic_num)
        self.atoms[oid] = a
        return a

    def add_bond(self, ida, idb, order=1, aromatic=False):
        a, b = self.atoms[ida], self.atoms[idb]
        bond = OBBond(a, b, order, aromatic)
        a.bonds.append(bond)
        b.bonds.append(bond)

    def GetAtomById(self, oid):
        return self.atoms[oid]


def OBAtomAtomIter(ob_atom):
    return iter([b.GetNbrAtom(ob_atom) for b in ob_atom.bonds])


def OBAtomBondIter(ob_atom):
    return iter(list(ob_atom.bonds))


def make_ob_module():
    ob = types.ModuleType('openbabel.openbabel')
    ob.OBAtomAtomIter = OBAtomAtomIter
    ob.OBAtomBondIter = OBAtomBondIter
    ob.OBMol = OBMol
    ob.Hydrogen = 1
    return ob


def install_stubs():
    """Register stand-in modules for Bio, openbabel and gemmi (idempotent)."""
    if 'Bio' in sys.modules and getattr(sys.modules['Bio'], '_arp_stub', False):
        return sys.modules['openbabel.openbabel']
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    bio = mod('Bio', _arp_stub=True)
    pdb = mod('Bio.PDB', NeighborSearch=NeighborSearch)
    bio.PDB = pdb
    pdb.Atom = mod('Bio.PDB.Atom', Atom=Atom, DisorderedAtom=DisorderedAtom)
    pdb.Residue = mod('Bio.PDB.Residue', Residue=Residue)
    pdb.PDBParser = mod('Bio.PDB.PDBParser', PDBParser=object)
    pdb.Polypeptide = mod('Bio.PDB.Polypeptide', PPBuilder=object)
    pdb.StructureBuilder = mod('Bio.PDB.StructureBuilder', StructureBuilder=object)
    mod('gemmi')
    ob = make_ob_module()
    sys.modules['openbabel.openbabel'] = ob
    mod('openbabel', openbabel=ob)
    return ob


# --------------------------------------------------------------------------
# seeded mini-complex generator
# --------------------------------------------------------------------------
# element -> (atomic number, vdw, cov); values in the style of OpenBabel's tables.  The contact
# code only ever sees them as per-atom Python floats (interactions.py:1501, :1509).
ELEMENTS = {
    'C': (6, 1.7, 0.76), 'N': (7, 1.55, 0.71), 'O': (8, 1.52, 0.66), 'S': (16, 1.8, 1.05),
    'H': (1, 1.1, 0.31), 'D': (1, 1.1, 0.31), 'ZN': (30, 1.39, 1.22), 'FE': (26, 2.05, 1.32),
    'CL': (17, 1.75, 1.02), 'F': (9, 1.47, 0.57), 'BR': (35, 1.85, 1.2), 'P': (15, 1.8, 1.07),
}
METALS = {'ZN', 'FE'}
HALOGENS = {'F', 'CL', 'BR', 'I', 'AT'}
RESNAMES = ('ALA', 'GLY', 'SER', 'LEU', 'PHE', 'ASP', 'LYS', 'MET', 'HIS', 'TYR', 'GLU', 'THR')


class MockStructure:
    def __init__(self):
        self.rings = collections.OrderedDict()
        self.amides = collections.OrderedDict()
        self.chains = []

    def get_chains(self):
        return iter(self.chains)

    def get_residues(self):
        for c in self.chains:
            yield from c.child_list

    def get_atoms(self):
        for r in self.get_residues():
            yield from r.child_list


class MockComplex:
    """Holds what InteractionComplex holds after initialize() (interactions.py:51-105, :288-327)."""

    def __init__(self):
        self.id = 'mock'
        self.biopython_str = MockStructure()
        self.s_atoms = []
        self.ob_mol = OBMol()
        self.ns = None
        self.ob_to_bio = {}
        self.bio_to_ob = {}
        self.component_types = {}
        self.selection = []
        self.selection_ring_ids = []
        self.selection_amide_ids = []
        self.polypeptide_residues = []
        self.selection_plus = []
        self.selection_plus_residues = []
        self.selection_plus_ring_ids = []
        self.selection_plus_amide_ids = []
        self.atom_contacts = []
        self.atom_plane_contacts = []
        self.plane_plane_contacts = []
        self.group_group_contacts = []
        self.group_plane_contacts = []


def residue_key(r):
    """A stable name of a residue across rebuilds of the same recipe."""
    return f'{r.get_parent().id}/{r.id[0].strip()}/{r.id[1]}/{r.id[2].strip()}/{r.resname}'


def flag_polypeptides(cx):
    """Polypeptide flags exactly as _handle_chains_residues_and_breaks leaves them
    (interactions.py:1663-1695): only residues of a polypeptide get prev_/next_residue."""
    cx.polypeptide_residues = set()
    for pp in cx.polypeptides:
        last_residue = None
        for residue in pp:
            cx.polypeptide_residues.add(residue)
            residue.is_polypeptide = True
            residue.prev_residue = last_residue
            residue.next_residue = None
            if last_residue:
                last_residue.next_residue = residue
            last_residue = residue


def _unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def _plane_normal(pts):
    """Mean of cross products of consecutive centre->vertex vectors (OBRing::findCenterAndNormal)."""
    c = pts.mean(0)
    n = np.zeros(3)
    for k in range(len(pts)):
        n += np.cross(pts[k] - c, pts[(k + 1) % len(pts)] - c)
    return c, n / np.linalg.norm(n)


def build_complex(seed=0, n_chains=2, n_res=24, n_waters=16, ligand=True, explicit_h=True,
                  degenerate=False, lone_xdonor=False, motifs=0, spread=1.0):
    """A small protein-like complex with every feature the contact rules look at.

    degenerate=True adds exact-geometry corner cases: coincident atoms, parallel and
    axis-aligned plane normals, a hydrogen sitting on its donor, atoms at exactly the cutoff.
    spread > 1 lets the backbone walk in a larger sphere (fewer steric clashes: protein-like density).
    motifs=m adds m copies each of the geometries the RARE rules need, every copy in residues of its own,
    anchored at least 3.4 A away from everything placed before: a C-Cl...O halogen bond (is_xbond, around its
    distance and angle thresholds), a C-H...Br-C contact (is_halogen_weak_hbond), a Lys-N+ ... Asp-O- pair (ionic,
    2.6-4.4 A), a zinc ion with three oxygen ligands (metal complex, 1.9-3.1 A) and a cluster of five waters
    (WATER_WATER entity class).
    """
    rng = np.random.default_rng(seed)
    cx = MockComplex()
    st = cx.biopython_str
    serial = [0]
    sphere = (4.2 * (n_chains * n_res) ** (1.0 / 3.0) + 3.0) * spread

    def add_atom(res, name, element, xyz, types=(), n_h=0):
        serial[0] += 1
        xyz = np.round(np.asarray(xyz, dtype=float), 3)
        a = Atom(res, name, element, xyz, serial[0])
        z, vdw, cov = ELEMENTS[element.strip().upper()]
        a.vdw_radius, a.cov_radius = float(vdw), float(cov)
        a.is_metal = element.upper() in METALS
        a.is_halogen = element.upper() in HALOGENS
        a.atom_types = set(types)
        for _ in range(n_h):
            a.h_coords.append(np.array(a.coord, dtype='d') + _unit(rng) * rng.uniform(0.95, 1.1))
        cx.s_atoms.append(a)
        cx.ob_mol.add_atom(serial[0], z)
        cx.ob_to_bio[serial[0]] = a
        cx.bio_to_ob[a] = serial[0]
        return a

    def bond(a, b, order=1, aromatic=False):
        cx.ob_mol.add_bond(a.serial_number, b.serial_number, order, aromatic)

    def add_ring(atoms, residue):
        e = len(st.rings)
        pts = np.array([np.array(a.coord, dtype='d') for a in atoms])
        c, n = _plane_normal(pts)
        st.rings[e] = {'ring_id': e, 'center': c, 'normal': n, 'normal_opp': -n, 'atoms': list(atoms),
                       'ob_atom_ids': [a.serial_number for a in atoms], 'residue': residue}
        return e

    def add_amide(n_atom, c_atom, o_atom, ca_atom):
        # interactions.py:1564-1589: float32 centre (C-N midpoint) and float32 SVD normal
        e = len(st.amides)
        bio = [n_atom, c_atom, o_atom, ca_atom]
        con = np.array([c_atom.coord, o_atom.coord, n_atom.coord])
        cn = np.array([c_atom.coord, n_atom.coord])
        centroid = con.sum(0) / float(len(con))
        bond_centroid = cn.sum(0) / float(len(cn))
        _, _, vh = np.linalg.svd(con - centroid)
        normal = np.array(vh.conj().transpose()[:, -1])
        residues = [x.get_parent() for x in bio]
        st.amides[e] = {'amide_id': e, 'center': bond_centroid, 'normal': normal, 'normal_opp': -normal,
                        'atoms': bio, 'residue': max(residues, key=residues.count)}
        return e

    side_types = [
        ('CG', 'C', {'hydrophobe', 'weak hbond donor'}, 2), ('CD', 'C', {'hydrophobe', 'weak hbond donor'}, 1),
        ('OG', 'O', {'hbond acceptor', 'hbond donor', 'xbond acceptor'}, 1),
        ('OD1', 'O', {'hbond acceptor', 'neg ionisable', 'xbond acceptor'}, 0),
        ('NZ', 'N', {'hbond donor', 'pos ionisable'}, 3), ('SD', 'S', {'hydrophobe', 'xbond acceptor', 'hbond acceptor'}, 0),
        ('NE2', 'N', {'hbond acceptor', 'hbond donor', 'aromatic', 'xbond acceptor'}, 1),
        ('OH', 'O', {'hbond acceptor', 'hbond donor', 'weak hbond acceptor'}, 1),
    ]

    all_pp = []
    for ci in range(n_chains):
        chain = Chain('AB'[ci % 2] + ('' if ci < 2 else str(ci)))
        st.chains.append(chain)
        pos = _unit(rng) * rng.uniform(0, sphere * 0.5)
        peptide = []
        prev_c = None
        for ri in range(n_res):
            step = _unit(rng) * 3.8
            if np.linalg.norm(pos + step) > sphere:
                step = -step
            pos = pos + step
            resname = RESNAMES[int(rng.integers(len(RESNAMES)))]
            res = Residue(chain, ' ', ri + 1, ' ', resname)
            ca = add_atom(res, 'CA', 'C', pos, {'weak hbond donor'}, 1)
            n = add_atom(res, 'N', 'N', pos + _unit(rng) * 1.46, {'hbond donor'}, 1)
            c = add_atom(res, 'C', 'C', pos + _unit(rng) * 1.52, {'carbonyl carbon'})
            o = add_atom(res, 'O', 'O', c.coord + _unit(rng) * 1.23, {'hbond acceptor', 'carbonyl oxygen', 'xbond acceptor'})
            cb = add_atom(res, 'CB', 'C', pos + _unit(rng) * 1.53, {'hydrophobe', 'weak hbond donor'}, 2)
            bond(n, ca); bond(ca, c); bond(c, o, 2); bond(ca, cb)
            if prev_c is not None:
                bond(prev_c, n)
            last = cb
            if resname in ('PHE', 'TYR', 'HIS'):
                centre = cb.coord + _unit(rng) * 2.9
                u = _unit(rng); v = np.cross(u, _unit(rng)); v /= np.linalg.norm(v)
                ring_atoms = []
                for k in range(6):
                    ang = 2 * np.pi * k / 6
                    ra = add_atom(res, f'CR{k}', 'C', centre + 1.39 * (np.cos(ang) * u + np.sin(ang) * v),
                                  {'aromatic', 'hydrophobe', 'weak hbond donor'}, 1)
                    ring_atoms.append(ra)
                for k in range(6):
                    bond(ring_atoms[k], ring_atoms[(k + 1) % 6], 1, True)
                bond(cb, ring_atoms[0])
                add_ring(ring_atoms, res)
            else:
                for _ in range(int(rng.integers(0, 4))):
                    nm, el, ty, nh = side_types[int(rng.integers(len(side_types)))]
                    if any(a.name == nm for a in res.child_list):
                        continue
                    sa = add_atom(res, nm, el, last.coord + _unit(rng) * 1.5, ty, nh)
                    bond(last, sa)
                    last = sa
            if resname == 'MET':
                sd = [a for a in res.child_list if a.name == 'SD']
                if not sd:
                    sa = add_atom(res, 'SD', 'S', last.coord + _unit(rng) * 1.8, {'hydrophobe', 'xbond acceptor'})
                    bond(last, sa)
            if explicit_h and rng.random() < 0.5:
                h = add_atom(res, 'H', 'H', n.coord + _unit(rng) * 1.0)
                bond(n, h)
            if prev_c is not None:
                pr = prev_c.get_parent()
                prev_n = [a for a in res.child_list if a.name == 'N'][0]
                pca = [a for a in pr.child_list if a.name == 'CA'][0]
                po = [a for a in pr.child_list if a.name == 'O'][0]
                add_amide(prev_n, prev_c, po, pca)
            prev_c = c
            peptide.append(res)
            # a chain break in the middle of the first chain: two polypeptides
            if ci == 0 and ri == n_res // 2:
                all_pp.append(peptide)
                peptide = []
                prev_c = None
        all_pp.append(peptide)

    cx.polypeptides = all_pp
    flag_polypeptides(cx)

    het = Chain('A') if not st.chains else st.chains[0]
    seq = 500
    centre_pts = [np.array(a.coord, dtype='d') for a in cx.s_atoms]

    def near():
        return centre_pts[int(rng.integers(len(centre_pts)))] + _unit(rng) * rng.uniform(2.4, 3.6)

    for _ in range(n_waters):
        seq += 1
        res = Residue(het, 'W', seq, ' ', 'HOH')
        add_atom(res, 'O', 'O', near(), {'hbond acceptor', 'hbond donor'}, 2)
    # a metal ion next to an acceptor
    seq += 1
    res = Residue(het, 'H_ZN', seq, ' ', 'ZN')
    acc = [a for a in cx.s_atoms if 'hbond acceptor' in a.atom_types]
    add_atom(res, 'ZN', 'ZN', np.array(acc[int(rng.integers(len(acc)))].coord, dtype='d') + _unit(rng) * 2.1)

    lig_atoms = []
    if ligand:
        seq += 1
        res = Residue(het, 'H_LIG', seq, ' ', 'LIG')
        base = near()
        u = _unit(rng); v = np.cross(u, _unit(rng)); v /= np.linalg.norm(v)
        ring_atoms = []
        for k in range(6):
            ang = 2 * np.pi * k / 6
            ring_atoms.append(add_atom(res, f'C{k+1}', 'C', base + 1.39 * (np.cos(ang) * u + np.sin(ang) * v),
                                       {'aromatic', 'hydrophobe', 'weak hbond donor'}, 1))
        for k in range(6):
            bond(ring_atoms[k], ring_atoms[(k + 1) % 6], 1, True)
        add_ring(ring_atoms, res)
        cl = add_atom(res, 'CL1', 'CL', ring_atoms[0].coord + u * 1.74, {'xbond donor', 'weak hbond acceptor', 'hydrophobe'})
        bond(ring_atoms[0], cl)
        br = add_atom(res, 'BR1', 'BR', ring_atoms[3].coord - u * 1.9, {'xbond donor', 'weak hbond acceptor'})
        bond(ring_atoms[3], br)
        if lone_xdonor:
            # an xbond donor without any bond: utils.is_xbond dereferences None (utils.py:173)
            xacc = [a for a in cx.s_atoms if 'xbond acceptor' in a.atom_types and a.get_parent() is not res]
            tgt = xacc[int(rng.integers(len(xacc)))]
            add_atom(res, 'F9', 'F', np.array(tgt.coord, dtype='d') + _unit(rng) * 2.6, {'weak hbond acceptor', 'xbond donor'})
        n1 = add_atom(res, 'N1', 'N', ring_atoms[1].coord + v * 1.4, {'hbond donor', 'pos ionisable'}, 2)
        bond(ring_atoms[1], n1)
        o1 = add_atom(res, 'O1', 'O', ring_atoms[4].coord - v * 1.4, {'hbond acceptor', 'neg ionisable', 'xbond acceptor'})
        bond(ring_atoms[4], o1, 2)
        cc = add_atom(res, 'C7', 'C', n1.coord + _unit(rng) * 1.35, {'carbonyl carbon'})
        oo = add_atom(res, 'O7', 'O', cc.coord + _unit(rng) * 1.23, {'hbond acceptor', 'carbonyl oxygen'})
        c8 = add_atom(res, 'C8', 'C', cc.coord + _unit(rng) * 1.5, {'hydrophobe', 'weak hbond donor'}, 3)
        bond(n1, cc); bond(cc, oo, 2); bond(cc, c8)
        add_amide(n1, cc, oo, c8)
        lig_atoms = list(res.child_list)
        # covalent link ligand -> protein (inter-residue OB bond; SIFt[1])
        prot = [a for a in cx.s_atoms if a.get_parent().id[0] == ' ' and a.element != 'H']
        d = [np.linalg.norm(np.array(a.coord, dtype='d') - np.array(c8.coord, dtype='d')) for a in prot]
        bond(c8, prot[int(np.argmin(d))])

    # a disulfide-like inter-residue bond between two non-adjacent residues
    sulf = [a for a in cx.s_atoms if a.element == 'S']
    if len(sulf) >= 2:
        bond(sulf[0], sulf[-1])

    if motifs:
        placed = np.zeros((len(cx.s_atoms) + 18 * motifs + 16, 3))
        n_placed = [len(cx.s_atoms)]
        placed[:n_placed[0]] = [np.array(a.coord, dtype='d') for a in cx.s_atoms]
        reach = sphere + 6.0

        def anchor(clear=3.4, room=4.5):
            """a point at least `clear` from every atom so far (the motif then occupies a ball of radius `room`)"""
            for _ in range(400):
                p = _unit(rng) * reach * rng.uniform(0.2, 1.0) ** (1.0 / 3.0)
                d = placed[:n_placed[0]] - p
                if np.sqrt((d * d).sum(axis=1).min()) >= clear + room * 0.5:
                    return p
            return _unit(rng) * (reach + rng.uniform(4.0, 30.0))

        def het_res(code):
            nonlocal seq
            seq += 1
            return Residue(het, 'H_' + code, seq, ' ', code)

        def put(res, name, element, xyz, types=(), n_h=0):
            a = add_atom(res, name, element, xyz, types, n_h)
            placed[n_placed[0]] = np.array(a.coord, dtype='d')
            n_placed[0] += 1
            return a

        def perp(u):
            v = np.cross(u, _unit(rng))
            return v / np.linalg.norm(v)

        for _ in range(motifs):
            # (1) C-Cl ... O: angle C-Cl-O between ~95 and 180 degrees, Cl...O between 2.9 and 3.6 A
            p = anchor(); u = _unit(rng)
            r1 = het_res('CLX')
            c1 = put(r1, 'C1', 'C', p, {'hydrophobe'})
            cl = put(r1, 'CL1', 'CL', p + u * 1.74, {'xbond donor', 'weak hbond acceptor', 'hydrophobe'})
            bond(c1, cl)
            d = u + perp(u) * rng.uniform(0.0, 1.1)
            d /= np.linalg.norm(d)
            put(het_res('ACX'), 'O1', 'O', np.array(cl.coord, dtype='d') + d * rng.uniform(2.9, 3.6),
                {'hbond acceptor', 'xbond acceptor'})
            # (2) C-H ... Br-C: the hydrogen 2.4-3.4 A from the bromine, C-Br...H between ~40 and 180 degrees
            p = anchor(); w = _unit(rng)
            r2 = het_res('BRX')
            c2 = put(r2, 'C1', 'C', p, {'hydrophobe'})
            br = put(r2, 'BR1', 'BR', p + w * 1.9, {'weak hbond acceptor'})
            bond(c2, br)
            d = w * rng.uniform(-0.2, 1.0) + perp(w) * rng.uniform(0.3, 1.0)
            d /= np.linalg.norm(d)
            hpos = np.array(br.coord, dtype='d') + d * rng.uniform(2.4, 3.4)
            don = put(het_res('WDN'), 'C1', 'C', hpos + d * 1.09 + _unit(rng) * 0.15, {'weak hbond donor', 'hydrophobe'})
            don.h_coords.append(hpos)
            # (3) N+ ... O-: 2.6-4.4 A (ionic needs <= 4.0)
            p = anchor()
            put(het_res('LYX'), 'NZ', 'N', p, {'hbond donor', 'pos ionisable'}, 3)
            put(het_res('ASX'), 'OD1', 'O', p + _unit(rng) * rng.uniform(2.6, 4.4), {'hbond acceptor', 'neg ionisable', 'xbond acceptor'})
            # (4) Zn with three oxygen ligands at 1.9-3.1 A (metal complex needs <= 2.8)
            p = anchor()
            put(het_res('ZN'), 'ZN', 'ZN', p)
            for k in range(3):
                put(het_res('ACM'), 'O1', 'O', p + _unit(rng) * rng.uniform(1.9, 3.1), {'hbond acceptor'})
            # (5) five waters within a 2.2 A ball
            p = anchor()
            for k in range(5):
                seq += 1
                put(Residue(het, 'W', seq, ' ', 'HOH'), 'O', 'O', p + _unit(rng) * rng.uniform(0.8, 2.2) * (k > 0),
                    {'hbond acceptor', 'hbond donor'}, 2)

    if degenerate:
        seq += 1
        res = Residue(het, 'H_DEG', seq, ' ', 'DEG')
        p0 = near()
        a0 = add_atom(res, 'X1', 'O', p0, {'hbond acceptor', 'hbond donor'}, 0)
        a0.h_coords.append(np.array(a0.coord, dtype='d'))            # hydrogen on the donor: NaN angle
        seq += 1
        res2 = Residue(het, 'H_DEG', seq, ' ', 'DEG')
        add_atom(res2, 'X2', 'N', a0.coord, {'hbond acceptor', 'hbond donor', 'xbond acceptor'}, 1)   # coincident atoms
        seq += 1
        res3 = Residue(het, 'H_DEG', seq, ' ', 'DEG')
        add_atom(res3, 'X3', 'C', np.array(a0.coord, dtype='d') + np.array([5.0, 0, 0]), {'hydrophobe'})  # exactly 5.0
        add_atom(res3, 'X4', 'C', np.array(a0.coord, dtype='d') + np.array([3.0, 4.0, 0]), {'hydrophobe'})  # 3-4-5
        # rings with exact normals: parallel, antiparallel, perpendicular, zero
        for k, (off, nrm) in enumerate((((0, 0, 0), (0, 0, 1)), ((0, 0, 3.5), (0, 0, 1)), ((0, 3.0, 0), (0, 0, -1)),
                                        ((3.0, 0, 0), (1, 0, 0)), ((0, 0, -3.0), (0, 0, 0)),
                                        ((1.0, 1.0, 3.0), (0.6, 0.0, 0.8)))):
            e = len(st.rings)
            owner = res3 if k % 2 else res2
            st.rings[e] = {'ring_id': e, 'center': np.array(p0, dtype='d') + np.array(off, dtype='d') + 8.0,
                           'normal': np.array(nrm, dtype='d'), 'normal_opp': -np.array(nrm, dtype='d'),
                           'atoms': list(owner.child_list), 'ob_atom_ids': [], 'residue': owner}
        for k, (off, nrm) in enumerate((((0, 0, 0), (0, 0, 1)), ((0, 0, 3.0), (0, 0, 1)), ((0, 2.0, 2.0), (0, 1, 0)),
                                        ((0, 0, -2.5), (0, 0, 0)))):
            e = len(st.amides)
            owner = res3 if k % 2 else res2
            st.amides[e] = {'amide_id': e, 'center': (np.array(p0) + np.array(off) + 8.0).astype('f'),
                            'normal': np.array(nrm, dtype='f'), 'normal_opp': -np.array(nrm, dtype='f'),
                            'atoms': list(owner.child_list), 'residue': owner}

    st.chains = list(dict.fromkeys(st.chains + [het]))
    for r in st.get_residues():
        cx.component_types[r.resname] = {' ': 'P', 'W': 'W'}.get(r.id[0], 'B')
    cx.lig_atoms = lig_atoms
    return cx
