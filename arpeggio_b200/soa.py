"""Structure-of-arrays containers handed across the C ABI.

``AtomSoA`` is the flat image of the reference's ``selection_plus`` atom list
(arpeggio/core/interactions.py:1426, :1442) with everything the contact loop
reads from the BioPython/OpenBabel objects (interactions.py:707-936); ``PlaneSoA``
is the image of ``structure.rings`` / ``structure.amides`` (interactions.py:1720-1725,
:1582-1589).  The containers own C-contiguous NumPy arrays of exactly the dtypes
``include/arpeggio_cuda.h`` declares, and hand out ctypes views of themselves.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import abi


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


@dataclass
class AtomSoA:
    xyz: np.ndarray                       # float32 [N,3]
    feat: np.ndarray                      # uint32 [N]
    res_id: np.ndarray                    # int32 [N]
    rad_class: np.ndarray                 # uint16 [N]
    vdw: np.ndarray                       # float64 [K]
    cov: np.ndarray                       # float64 [K]
    res_prev: np.ndarray                  # int32 [Rs]
    res_next: np.ndarray                  # int32 [Rs]
    res_flags: np.ndarray                 # uint8 [Rs]
    bond_off: Optional[np.ndarray] = None  # int32 [N+1]
    bond_nbr: Optional[np.ndarray] = None  # int32 [E]
    h_off: Optional[np.ndarray] = None     # int32 [N+1]
    h_xyz: Optional[np.ndarray] = None     # float64 [H,3]
    xnbr_xyz: Optional[np.ndarray] = None  # float32 [N,3]
    struct_off: Optional[np.ndarray] = None  # int32 [S+1]
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        self.xyz = _c(self.xyz, np.float32).reshape(-1, 3)
        n = self.xyz.shape[0]
        self.feat = _c(self.feat, np.uint32)
        self.res_id = _c(self.res_id, np.int32)
        self.rad_class = _c(self.rad_class, np.uint16)
        self.vdw = _c(self.vdw, np.float64)
        self.cov = _c(self.cov, np.float64)
        self.res_prev = _c(self.res_prev, np.int32)
        self.res_next = _c(self.res_next, np.int32)
        self.res_flags = _c(self.res_flags, np.uint8)
        if self.bond_off is not None:
            self.bond_off = _c(self.bond_off, np.int32)
            self.bond_nbr = _c(self.bond_nbr if self.bond_nbr is not None else [], np.int32)
        if self.h_off is not None:
            self.h_off = _c(self.h_off, np.int32)
            self.h_xyz = _c(self.h_xyz if self.h_xyz is not None else np.zeros((0, 3)), np.float64).reshape(-1, 3)
        if self.xnbr_xyz is not None:
            self.xnbr_xyz = _c(self.xnbr_xyz, np.float32).reshape(-1, 3)
        if self.struct_off is not None:
            self.struct_off = _c(self.struct_off, np.int32)
        self.validate(n)

    # ------------------------------------------------------------------
    @property
    def n_atoms(self):
        return self.xyz.shape[0]

    @property
    def n_residues(self):
        return self.res_flags.shape[0]

    @property
    def n_structures(self):
        return 1 if self.struct_off is None else self.struct_off.shape[0] - 1

    def validate(self, n=None):
        n = self.n_atoms if n is None else n
        rs = self.res_flags.shape[0]
        k = self.vdw.shape[0]
        if not (self.feat.shape == (n,) and self.res_id.shape == (n,) and self.rad_class.shape == (n,)):
            raise ValueError('feat/res_id/rad_class must have one entry per atom')
        if self.cov.shape != (k,):
            raise ValueError('vdw and cov tables differ in length')
        if self.res_prev.shape != (rs,) or self.res_next.shape != (rs,):
            raise ValueError('res_prev/res_next/res_flags differ in length')
        if n:
            if self.res_id.min() < 0 or self.res_id.max() >= rs:
                raise ValueError('res_id out of range')
            if self.rad_class.max() >= k:
                raise ValueError('rad_class out of range')
        if rs and (self.res_prev.min() < -1 or self.res_prev.max() >= rs or
                   self.res_next.min() < -1 or self.res_next.max() >= rs):
            raise ValueError('res_prev/res_next out of range')
        for off, dat, name in ((self.bond_off, self.bond_nbr, 'bond'), (self.h_off, self.h_xyz, 'h')):
            if off is None:
                continue
            if off.shape != (n + 1,) or off[0] != 0 or np.any(np.diff(off) < 0) or off[-1] != dat.shape[0]:
                raise ValueError(f'{name}_off is not a CSR offset array over the atoms')
        if self.bond_off is not None and self.bond_nbr.size and (self.bond_nbr.min() < 0 or self.bond_nbr.max() >= n):
            raise ValueError('bond_nbr out of range')
        if self.h_off is not None and n and np.diff(self.h_off).max() > 255:
            raise ValueError('more than 255 hydrogens on one atom')
        if self.xnbr_xyz is not None and self.xnbr_xyz.shape != (n, 3):
            raise ValueError('xnbr_xyz must be [N,3]')
        if self.struct_off is not None:
            so = self.struct_off
            if so.shape[0] < 2 or so[0] != 0 or so[-1] != n or np.any(np.diff(so) < 0):
                raise ValueError('struct_off must partition the atoms')

    def input_bytes(self):
        """Algorithmic input bytes: the sum of the array sizes."""
        tot = 0
        for a in (self.xyz, self.feat, self.res_id, self.rad_class, self.vdw, self.cov, self.res_prev,
                  self.res_next, self.res_flags, self.bond_off, self.bond_nbr, self.h_off, self.h_xyz,
                  self.xnbr_xyz, self.struct_off):
            if a is not None:
                tot += a.nbytes
        return tot

    def as_ctypes(self):
        """The arp_atoms image of the arrays; cached for as long as the same array objects are in place (a batch
        driver asks for it once per upload, and building it costs about as much as a small structure's kernels)."""
        names = ('xyz', 'feat', 'res_id', 'rad_class', 'vdw', 'cov', 'res_prev', 'res_next',
                 'res_flags', 'bond_off', 'bond_nbr', 'h_off', 'h_xyz', 'xnbr_xyz', 'struct_off')
        arrays = tuple(getattr(self, n) for n in names)       # kept with the cache: their ids cannot be reused meanwhile
        key = tuple(id(a) for a in arrays)
        cached = self.__dict__.get('_ct_cache')
        if cached is not None and cached[0] == key:
            return cached[1]
        s = abi.ArpAtoms()
        s.n_atoms = self.n_atoms
        s.n_residues = self.n_residues
        s.n_rad_classes = self.vdw.shape[0]
        s.n_structures = self.n_structures
        for name in ('xyz', 'feat', 'res_id', 'rad_class', 'vdw', 'cov', 'res_prev', 'res_next',
                     'res_flags', 'bond_off', 'bond_nbr', 'h_off', 'h_xyz', 'xnbr_xyz', 'struct_off'):
            setattr(s, name, abi.ptr(getattr(self, name)))
        self.__dict__['_ct_cache'] = (key, s, arrays)
        return s

    def to_wire(self, h_decimals=3):
        """The same structure in its wire form (WireAtoms): fewer bytes over PCIe, decoded on the device."""
        return WireAtoms.from_soa(self, h_decimals)

    def structure(self, s):
        """The s-th structure of a batch as a stand-alone AtomSoA (residue ids re-based)."""
        if self.struct_off is None:
            if s != 0:
                raise IndexError(s)
            return self
        lo, hi = int(self.struct_off[s]), int(self.struct_off[s + 1])
        rid = self.res_id[lo:hi]
        r0, r1 = (int(rid.min()), int(rid.max()) + 1) if hi > lo else (0, 0)
        fix = lambda a: np.where(a >= 0, a - r0, -1)
        kw = dict(xyz=self.xyz[lo:hi], feat=self.feat[lo:hi], res_id=rid - r0, rad_class=self.rad_class[lo:hi],
                  vdw=self.vdw, cov=self.cov, res_prev=fix(self.res_prev[r0:r1]), res_next=fix(self.res_next[r0:r1]),
                  res_flags=self.res_flags[r0:r1])
        if self.bond_off is not None:
            b0, b1 = int(self.bond_off[lo]), int(self.bond_off[hi])
            kw.update(bond_off=self.bond_off[lo:hi + 1] - b0, bond_nbr=self.bond_nbr[b0:b1] - lo)
        if self.h_off is not None:
            h0, h1 = int(self.h_off[lo]), int(self.h_off[hi])
            kw.update(h_off=self.h_off[lo:hi + 1] - h0, h_xyz=self.h_xyz[h0:h1])
        if self.xnbr_xyz is not None:
            kw.update(xnbr_xyz=self.xnbr_xyz[lo:hi])
        return AtomSoA(**kw)

    @staticmethod
    def concat(parts):
        """Concatenate independent structures into one batch (no pair spans two structures)."""
        parts = list(parts)
        if not parts:
            raise ValueError('empty batch')
        vdw, cov = parts[0].vdw, parts[0].cov
        for p in parts[1:]:
            if not (np.array_equal(p.vdw, vdw) and np.array_equal(p.cov, cov)):
                raise ValueError('structures of one batch must share the radius tables')
            if p.struct_off is not None:
                raise ValueError('nested batches are not supported')
        n_off = np.cumsum([0] + [p.n_atoms for p in parts])
        r_off = np.cumsum([0] + [p.n_residues for p in parts])
        shift = lambda a, o: np.where(a >= 0, a + o, -1)
        have_b = any(p.bond_off is not None for p in parts)
        have_h = any(p.h_off is not None for p in parts)
        have_x = any(p.xnbr_xyz is not None for p in parts)
        kw = dict(
            xyz=np.concatenate([p.xyz for p in parts]),
            feat=np.concatenate([p.feat for p in parts]),
            res_id=np.concatenate([p.res_id + r_off[i] for i, p in enumerate(parts)]),
            rad_class=np.concatenate([p.rad_class for p in parts]),
            vdw=vdw, cov=cov,
            res_prev=np.concatenate([shift(p.res_prev, r_off[i]) for i, p in enumerate(parts)]),
            res_next=np.concatenate([shift(p.res_next, r_off[i]) for i, p in enumerate(parts)]),
            res_flags=np.concatenate([p.res_flags for p in parts]),
            struct_off=n_off.astype(np.int32))
        if have_b:
            cnt = np.concatenate([np.diff(p.bond_off) if p.bond_off is not None else np.zeros(p.n_atoms, np.int32) for p in parts])
            kw['bond_off'] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
            kw['bond_nbr'] = np.concatenate([p.bond_nbr + n_off[i] if p.bond_off is not None else np.zeros(0, np.int32)
                                             for i, p in enumerate(parts)])
        if have_h:
            cnt = np.concatenate([np.diff(p.h_off) if p.h_off is not None else np.zeros(p.n_atoms, np.int32) for p in parts])
            kw['h_off'] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
            kw['h_xyz'] = np.concatenate([p.h_xyz if p.h_off is not None else np.zeros((0, 3)) for p in parts])
        if have_x:
            kw['xnbr_xyz'] = np.concatenate([p.xnbr_xyz if p.xnbr_xyz is not None else np.zeros((p.n_atoms, 3), np.float32)
                                             for p in parts])
        return AtomSoA(**kw)


@dataclass
class PlaneSoA:
    center: np.ndarray      # [n,3] float64 (rings) or float32 (amides)
    normal: np.ndarray      # [n,3] same dtype as center
    res_id: np.ndarray      # int32 [n]
    flags: np.ndarray       # uint32 [n]
    is_f32: bool = False

    def __post_init__(self):
        dt = np.float32 if self.is_f32 else np.float64
        self.center = _c(self.center, dt).reshape(-1, 3)
        self.normal = _c(self.normal, dt).reshape(-1, 3)
        self.res_id = _c(self.res_id, np.int32)
        self.flags = _c(self.flags, np.uint32)
        n = self.center.shape[0]
        if self.normal.shape != (n, 3) or self.res_id.shape != (n,) or self.flags.shape != (n,):
            raise ValueError('plane arrays differ in length')

    @property
    def n(self):
        return self.center.shape[0]

    def input_bytes(self):
        return self.center.nbytes + self.normal.nbytes + self.res_id.nbytes + self.flags.nbytes

    def as_ctypes(self):
        s = abi.ArpPlanes()
        s.n = self.n
        s.is_f32 = 1 if self.is_f32 else 0
        s.center = abi.ptr(self.center)
        s.normal = abi.ptr(self.normal)
        s.res_id = abi.ptr(self.res_id)
        s.flags = abi.ptr(self.flags)
        return s

    @staticmethod
    def empty(is_f32=False):
        return PlaneSoA(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.int32), np.zeros(0, np.uint32), is_f32)


class WireAtoms:
    """An AtomSoA in the wire forms of arp_atoms (include/arpeggio_cuda.h): per-atom uint8 counts in place of the two
    int32 CSR offset arrays, the single-bond neighbours of the halogens (get_single_bond_neighbour, utils.py:163-176) as
    (atom index, coordinate) rows in place of a dense [N, 3] array, and -- when every hydrogen coordinate is a decimal
    fraction of `h_decimals` digits, as coordinates read from PDB / mmCIF text are -- the float64 hydrogen coordinates as
    int32 fixed point.  Every form is lossless (the fixed-point one is checked value by value and skipped otherwise); the
    library decodes them on the device after the copy, so results are those of the AtomSoA.  Accepted wherever an AtomSoA
    is uploaded (ContactEngine.upload_atoms / upload_atoms_batch, BatchRunner.run, engine.pinned_soa)."""

    NAMES = ('xyz', 'feat', 'res_id', 'rad_class', 'vdw', 'cov', 'res_prev', 'res_next', 'res_flags', 'bond_off', 'bond_cnt',
             'bond_nbr', 'h_off', 'h_cnt', 'h_xyz', 'h_fix', 'xnbr_idx', 'xnbr_xyz', 'struct_off')

    def __init__(self, h_fix_scale=0.0, _keep=None, **arrays):
        for k in self.NAMES:
            setattr(self, k, arrays.pop(k, None))
        if arrays:
            raise TypeError(f'unknown arrays {sorted(arrays)}')
        self.h_fix_scale = float(h_fix_scale)
        self._keep = _keep or []
        self._ct = None

    @classmethod
    def from_soa(cls, soa, h_decimals=3):
        n = soa.n_atoms
        kw = dict(xyz=soa.xyz, feat=soa.feat, res_id=soa.res_id, rad_class=soa.rad_class, vdw=soa.vdw, cov=soa.cov,
                  res_prev=soa.res_prev, res_next=soa.res_next, res_flags=soa.res_flags, struct_off=soa.struct_off)
        scale = 0.0
        if soa.bond_off is not None:
            cnt = np.diff(soa.bond_off)
            if n and cnt.max() > 255:
                kw.update(bond_off=soa.bond_off)
            else:
                kw.update(bond_cnt=cnt.astype(np.uint8))
            kw.update(bond_nbr=soa.bond_nbr)
        if soa.h_off is not None:
            kw.update(h_cnt=np.diff(soa.h_off).astype(np.uint8))          # AtomSoA.validate: at most 255 hydrogens per atom
            fix = None
            if h_decimals is not None and soa.h_xyz.size:
                s = float(10 ** int(h_decimals))
                q = np.rint(soa.h_xyz * s)
                if np.all(np.abs(q) < 2.0 ** 31) and np.array_equal(q / s, soa.h_xyz):      # value by value, NaN fails
                    fix, scale = q.astype(np.int32), s
            if fix is not None:
                kw.update(h_fix=fix)
            else:
                kw.update(h_xyz=soa.h_xyz)
        if soa.xnbr_xyz is not None:
            idx = np.flatnonzero(soa.feat & np.uint32(abi.F_HAS_XNBR)).astype(np.int32)
            kw.update(xnbr_idx=idx, xnbr_xyz=np.ascontiguousarray(soa.xnbr_xyz[idx]))
        return cls(h_fix_scale=scale, **kw)

    @property
    def n_atoms(self):
        return self.xyz.shape[0]

    @property
    def n_residues(self):
        return self.res_flags.shape[0]

    @property
    def n_structures(self):
        return 1 if self.struct_off is None else self.struct_off.shape[0] - 1

    def input_bytes(self):
        return sum(a.nbytes for a in (getattr(self, k) for k in self.NAMES) if a is not None)

    def as_ctypes(self):
        if self._ct is None:
            s = abi.ArpAtoms()
            s.n_atoms, s.n_residues, s.n_rad_classes, s.n_structures = self.n_atoms, self.n_residues, self.vdw.shape[0], self.n_structures
            for k in self.NAMES:
                setattr(s, k, abi.ptr(getattr(self, k)))
            s.h_fix_scale = self.h_fix_scale
            s.n_bond_nbr = 0 if self.bond_nbr is None else self.bond_nbr.shape[0]
            n_h = self.h_fix if self.h_fix is not None else self.h_xyz
            s.n_h = 0 if n_h is None else n_h.shape[0]
            s.n_xnbr = 0 if self.xnbr_idx is None else self.xnbr_idx.shape[0]
            self._ct = s
        return self._ct
