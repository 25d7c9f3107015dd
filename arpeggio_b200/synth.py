"""Seeded synthetic inputs of the benchmark configurations (BASELINE.json `configs`).

These are inputs only -- random atom clouds laid out as the SoA of include/arpeggio_cuda.h.
Nothing here computes contacts.
"""
import numpy as np

from . import abi
from .soa import AtomSoA, PlaneSoA

DENSITY = 0.05          # atoms per cubic Angstrom (heavy-atom density of a protein)

# radius classes: (vdw, cov) in the spirit of OpenBabel's element table: C N O S metal halogen
RADII = np.array([(1.70, 0.76), (1.55, 0.71), (1.52, 0.66), (1.80, 1.05), (2.05, 1.32), (1.75, 1.02)])
CLASS_P = np.array([0.55, 0.17, 0.20, 0.03, 0.02, 0.03])
C_CLASS, N_CLASS, O_CLASS, S_CLASS, METAL_CLASS, HALOGEN_CLASS = range(6)


def box_edge(n, density=DENSITY):
    return (n / density) ** (1.0 / 3.0)


def cloud_coords(rng, n, density=DENSITY):
    """Uniform cloud in a cube, rounded to 3 decimals like mmCIF Cartn_x/y/z, then float32."""
    edge = box_edge(max(n, 1), density)
    return np.round(rng.uniform(0.0, edge, size=(n, 3)), 3).astype(np.float32)


def cloud_uniform(n=10_000, seed=1):
    """configs[1]: uniform VdW radii, no features, every atom its own residue -> distance + bits 0..4."""
    rng = np.random.default_rng(seed)
    return AtomSoA(xyz=cloud_coords(rng, n), feat=np.zeros(n, np.uint32), res_id=np.arange(n, dtype=np.int32),
                   rad_class=np.zeros(n, np.uint16), vdw=np.array([1.70]), cov=np.array([0.76]),
                   res_prev=np.full(n, -1, np.int32), res_next=np.full(n, -1, np.int32),
                   res_flags=np.zeros(n, np.uint8))


def cloud_featured(n=100_000, seed=2, atoms_per_residue=8, chain_len=300, bonds=True, h_decimals=None):
    """configs[2]: 100k-atom cloud with SIFt feature masks, residues, hydrogens, bonds, halogen neighbours."""
    rng = np.random.default_rng(seed)
    xyz = cloud_coords(rng, n)
    rad_class = rng.choice(6, size=n, p=CLASS_P).astype(np.uint16)

    def bit(p):
        return rng.random(n) < p

    feat = np.zeros(n, np.uint32)
    for flag, p in ((abi.F_HBOND_ACCEPTOR, .25), (abi.F_HBOND_DONOR, .20), (abi.F_WEAK_HBOND_ACCEPTOR, .10),
                    (abi.F_WEAK_HBOND_DONOR, .45), (abi.F_XBOND_ACCEPTOR, .25), (abi.F_POS_IONISABLE, .03),
                    (abi.F_NEG_IONISABLE, .03), (abi.F_HYDROPHOBE, .35), (abi.F_CARBONYL_OXYGEN, .08),
                    (abi.F_CARBONYL_CARBON, .08), (abi.F_AROMATIC, .10), (abi.F_IS_WATER, .05),
                    (abi.F_IN_SELECTION, .10)):
        feat[bit(p)] |= flag
    halogen = rad_class == HALOGEN_CLASS
    feat[halogen] |= abi.F_IS_HALOGEN | abi.F_HAS_XNBR
    feat[halogen & bit(1 / 3)] |= abi.F_XBOND_DONOR          # ~1 % of the atoms
    feat[halogen & bit(0.5)] |= abi.F_WEAK_HBOND_ACCEPTOR
    feat[rad_class == METAL_CLASS] |= abi.F_IS_METAL
    feat[rad_class == C_CLASS] |= abi.F_ELEM_C
    feat[(rad_class == S_CLASS) & bit(0.3)] |= abi.F_MET_SULPHUR
    # waters carry both hbond types (interactions.py:1962-1964)
    water = (feat & abi.F_IS_WATER) != 0
    feat[water] |= abi.F_HBOND_ACCEPTOR | abi.F_HBOND_DONOR

    # residues: consecutive runs of atoms; polypeptide chains of chain_len residues
    res_id = (np.arange(n) // atoms_per_residue).astype(np.int32)
    rs = int(res_id[-1]) + 1 if n else 0
    r = np.arange(rs)
    res_prev = np.where(r % chain_len == 0, -1, r - 1).astype(np.int32)
    res_next = np.where((r % chain_len == chain_len - 1) | (r == rs - 1), -1, r + 1).astype(np.int32)
    res_flags = np.zeros(rs, np.uint8)
    pp = rng.random(rs) < 0.9
    res_flags[pp] = abi.R_IS_POLYPEPTIDE | abi.R_HAS_LINKS
    res_prev[~pp] = -1
    res_next[~pp] = -1

    # hydrogens: donors and weak donors carry 1..3 H at 1.0 A in random directions (float64)
    donors = (feat & (abi.F_HBOND_DONOR | abi.F_WEAK_HBOND_DONOR)) != 0
    h_cnt = np.where(donors, rng.integers(1, 4, size=n), 0).astype(np.int32)
    h_off = np.concatenate([[0], np.cumsum(h_cnt)]).astype(np.int32)
    owner = np.repeat(np.arange(n), h_cnt)
    d = rng.normal(size=(owner.shape[0], 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    h_xyz = xyz[owner].astype(np.float64) + d
    if h_decimals is not None:          # as read from PDB / mmCIF text (3 decimals): the wire form can carry these as int32 fixed point
        h_xyz = np.round(h_xyz, h_decimals)

    # halogen / xbond-donor single-bond neighbour at 1.8 A
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xnbr = (xyz.astype(np.float64) + 1.8 * d).astype(np.float32)
    xnbr[~halogen] = 0

    kw = {}
    if bonds and n > 1:
        # 2 % of the atoms are covalently bound to one atom of another residue within 2 A
        from scipy.spatial import cKDTree
        tree = cKDTree(xyz.astype(np.float64))
        cand = np.nonzero(bit(0.02))[0]
        nbrs = tree.query_ball_point(xyz[cand].astype(np.float64), 2.0)
        pairs = set()
        for i, lst in zip(cand, nbrs):
            for j in lst:
                if res_id[j] != res_id[i]:
                    pairs.add((min(i, j), max(i, j)))
                    break
        if pairs:
            e = np.array(sorted(pairs), dtype=np.int64)
            src = np.concatenate([e[:, 0], e[:, 1]])
            dst = np.concatenate([e[:, 1], e[:, 0]])
            order = np.lexsort((dst, src))
            src, dst = src[order], dst[order]
        else:
            src = dst = np.zeros(0, np.int64)
        kw['bond_off'] = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=n))]).astype(np.int32)
        kw['bond_nbr'] = dst.astype(np.int32)

    return AtomSoA(xyz=xyz, feat=feat, res_id=res_id, rad_class=rad_class, vdw=RADII[:, 0].copy(), cov=RADII[:, 1].copy(),
                   res_prev=res_prev, res_next=res_next, res_flags=res_flags, h_off=h_off, h_xyz=h_xyz, xnbr_xyz=xnbr, **kw)


def structure_batch(n_structures=1024, atoms=20_000, seed0=1000, first=0, bonds=False):
    """configs[4]: independent featured structures [first, first + n_structures) of the batch, concatenated."""
    return AtomSoA.concat([cloud_featured(atoms, seed0 + first + s, bonds=bonds) for s in range(n_structures)])


def plane_set(n_rings=2048, n_amides=12_500, n_atoms=100_000, seed=3, n_residues=None):
    """configs[3]: synthetic ring (float64) and amide (float32) planes in the box of the n_atoms cloud."""
    rng = np.random.default_rng(seed)
    edge = box_edge(n_atoms)
    n_residues = n_residues or max(n_atoms // 8, 1)

    def planes(m, is_f32):
        dt = np.float32 if is_f32 else np.float64
        center = rng.uniform(0.0, edge, size=(m, 3)).astype(dt)
        normal = rng.normal(size=(m, 3))
        normal /= np.linalg.norm(normal, axis=1, keepdims=True)
        normal = normal.astype(dt)
        flags = np.full(m, abi.P_IN_SELECTION_PLUS, np.uint32)
        flags[rng.random(m) < 0.1] |= abi.P_IN_SELECTION
        flags[rng.random(m) < 0.05] = 0                      # outside the binding site
        return PlaneSoA(center, normal, rng.integers(0, n_residues, size=m).astype(np.int32), flags, is_f32)

    return planes(n_rings, False), planes(n_amides, True)
