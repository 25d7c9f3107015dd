"""``ContactEngine``: one CUDA context (device + stream) of libarpeggio_cuda.so.

Thin, explicit wrapper of the C ABI (include/arpeggio_cuda.h).  It owns no algorithm: uploads a
structure-of-arrays image of the reference's ``selection_plus`` atom list, runs the kernels that
replace ``InteractionComplex._calculate_atom_contacts`` / ``_calculate_ring_contacts`` /
``_calculate_group_contacts`` (arpeggio/core/interactions.py:693-936, :938-1194, :1208-1382), and
returns the record streams as NumPy structured arrays.  A context is not re-entrant; use one per
host thread.  ctypes releases the GIL during the calls.
"""
import ctypes as C

import numpy as np

from . import abi, params as arp_params
from ._lib import ArpeggioCudaError, lib


def device_count():
    n = lib().arp_device_count()
    if n < 0:
        raise ArpeggioCudaError(n, 'cudaGetDeviceCount failed')
    return n


class PinnedBuffer:
    """Page-locked host memory from arp_host_alloc, exposed as a NumPy array."""

    def __init__(self, nbytes):
        self._ptr = C.c_void_p()
        rc = lib().arp_host_alloc(C.byref(self._ptr), int(nbytes))
        if rc != abi.OK:
            raise ArpeggioCudaError(rc, 'pinned host allocation failed')
        self.nbytes = int(nbytes)

    def array(self, dtype, count=None):
        dtype = np.dtype(dtype)
        count = self.nbytes // dtype.itemsize if count is None else int(count)
        if count * dtype.itemsize > self.nbytes:
            raise ValueError('pinned buffer too small')
        if count == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (count * dtype.itemsize)).from_address(self._ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def free(self):
        if self._ptr:
            lib().arp_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_soa(soa):
    """A copy of an AtomSoA in ONE block of page-locked host memory (arp_host_alloc), every array at a 256-byte
    offset: arp_upload_atoms then moves the whole structure with a single asynchronous DMA (one transfer instead of
    fourteen) and the device arrays are views of one arena.  The returned object keeps the block alive (``_keep``)."""
    from .soa import AtomSoA, WireAtoms
    wire = isinstance(soa, WireAtoms)
    names = WireAtoms.NAMES if wire else ('xyz', 'feat', 'res_id', 'rad_class', 'vdw', 'cov', 'res_prev', 'res_next', 'res_flags',
                                          'bond_off', 'bond_nbr', 'h_off', 'h_xyz', 'xnbr_xyz', 'struct_off')
    arrays = [(k, getattr(soa, k, None)) for k in names]
    offsets, total = {}, 0
    for k, a in arrays:
        if a is not None:
            offsets[k] = total
            total += (a.nbytes + 255) // 256 * 256
    block = PinnedBuffer(max(total, 256))
    raw = block.array(np.uint8)
    fields = {}
    for k, a in arrays:
        if a is None:
            fields[k] = None
            continue
        v = raw[offsets[k]:offsets[k] + a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        fields[k] = v
    if wire:
        return WireAtoms(h_fix_scale=soa.h_fix_scale, _keep=[block], **fields)
    return AtomSoA(**fields, _keep=[block])


class CompactPairs:
    """Compact view of an (i, j)-sorted record stream: the records of bgn atom i are rec[row_off[i]:row_off[i + 1]]
    (fields j, mask); dist (float32, same order) is optional.  to_records() gives the 16-byte PAIR_DTYPE array."""

    def __init__(self, row_off, rec_buf, dist_buf):
        self.row_off, self.rec_buf, self.dist_buf = row_off, rec_buf, dist_buf
        self.rec, self.dist, self.n, self.n_atoms, self.atom_base = rec_buf[:0], None, 0, 0, 0

    def view(self, n_atoms, n, with_dist):
        v = CompactPairs(self.row_off, self.rec_buf, self.dist_buf)
        v.n_atoms, v.n = n_atoms, n
        v.row_off = self.row_off[:n_atoms + 1]
        v.rec = self.rec_buf[:n]
        v.dist = self.dist_buf[:n] if with_dist and self.dist_buf is not None else None
        return v

    @property
    def nbytes(self):
        return self.row_off.nbytes + self.rec.nbytes + (self.dist.nbytes if self.dist is not None else 0)

    def structure(self, a0, a1):
        """The rows of atoms [a0, a1) -- one structure of a batch -- as a CompactPairs of its own (views; the j of its
        records stay GLOBAL atom indices: subtract a0 for indices local to the structure)."""
        v = CompactPairs(self.row_off, self.rec_buf, self.dist_buf)
        r0, r1 = int(self.row_off[a0]), int(self.row_off[a1])
        v.n_atoms, v.n, v.atom_base = a1 - a0, r1 - r0, a0
        v.row_off = self.row_off[a0:a1 + 1] - np.uint32(r0)
        v.rec = self.rec[r0:r1]
        v.dist = self.dist[r0:r1] if self.dist is not None else None
        return v

    def to_records(self, dist=None):
        """PAIR_DTYPE[n] through the library's host unpacker (arp_pairs_unpack)."""
        out = np.empty(self.n, dtype=abi.PAIR_DTYPE)
        d = self.dist if dist is None else np.ascontiguousarray(dist, np.float32)
        rc = lib().arp_pairs_unpack(self.row_off.ctypes.data, self.rec.ctypes.data if self.n else None,
                                    d.ctypes.data if d is not None and self.n else None, self.n_atoms,
                                    out.ctypes.data if self.n else None, self.n)
        if rc != abi.OK:
            raise ArpeggioCudaError(rc, 'arp_pairs_unpack failed')
        return out


class PackedPairs:
    """Packed view of an (i, j)-sorted record stream (arp_pairs_fetch_packed): the records of bgn atom i are words
    row_off[i] .. row_off[i + 1]; a word holds j in its low bits_j bits and the 15 SIFt bits above them (lo32: bits
    0..31, hi8: bits 32..39, only for more than 131072 atoms).  4 (or 5) bytes per record + 4 per atom.  The entity
    class is recomputed from the atoms' feat words by to_records(); n_faults counts records whose xbond-without-neighbour
    fault bit is not in the words."""

    def __init__(self, row_off, lo_buf, hi_buf, dist_buf):
        self.row_off, self.lo_buf, self.hi_buf, self.dist_buf = row_off, lo_buf, hi_buf, dist_buf
        self.lo, self.hi, self.dist, self.n, self.n_atoms, self.bits_j, self.n_faults, self.atom_base = lo_buf[:0], None, None, 0, 0, 1, 0, 0

    def view(self, n_atoms, n, bits_j, n_faults, with_dist):
        v = PackedPairs(self.row_off, self.lo_buf, self.hi_buf, self.dist_buf)
        v.n_atoms, v.n, v.bits_j, v.n_faults = n_atoms, n, bits_j, n_faults
        v.row_off = self.row_off[:n_atoms + 1]
        v.lo = self.lo_buf[:n]
        v.hi = self.hi_buf[:n] if bits_j + 15 > 32 and self.hi_buf is not None else None
        v.dist = self.dist_buf[:n] if with_dist and self.dist_buf is not None else None
        return v

    def structure(self, lo, hi):
        """The rows of atoms lo..hi-1 (one structure of a batch) as a PackedPairs of their own: row offsets rebased (a small
        copy), words and distances as views.  The words of a batch hold structure-local j, so to_records(feat of that
        structure) gives indices local to the structure.  n_faults stays the count of the whole batch."""
        r0, r1 = int(self.row_off[lo]), int(self.row_off[hi])
        v = PackedPairs(self.row_off, self.lo_buf, self.hi_buf, self.dist_buf)
        v.row_off = self.row_off[lo:hi + 1] - np.uint32(r0)
        v.lo = self.lo[r0:r1]
        v.hi = self.hi[r0:r1] if self.hi is not None else None
        v.dist = self.dist[r0:r1] if self.dist is not None else None
        v.n, v.n_atoms, v.bits_j, v.n_faults, v.atom_base = r1 - r0, hi - lo, self.bits_j, self.n_faults, self.atom_base + lo
        return v

    @property
    def nbytes(self):
        return self.row_off.nbytes + self.lo.nbytes + (self.hi.nbytes if self.hi is not None else 0) + (self.dist.nbytes if self.dist is not None else 0)

    def to_records(self, feat, dist=None, struct_off=None, threads=0):
        """PAIR_DTYPE[n] through the library's host unpacker; feat: the uploaded atoms' feature words (uint32[n_atoms]).
        struct_off: the atom offsets of a batch (the words of a batch hold structure-local j; the records come out with
        the batch-global indices).  A view made by structure() needs none: its records are local to the structure.
        threads > 1: host threads that share the rows."""
        out = np.empty(self.n, dtype=abi.PAIR_DTYPE)
        feat = np.ascontiguousarray(feat, np.uint32)
        d = self.dist if dist is None else np.ascontiguousarray(dist, np.float32)
        so = None if struct_off is None else np.ascontiguousarray(struct_off, np.int32)
        rc = lib().arp_pairs_unpack_packed(self.row_off.ctypes.data, self.lo.ctypes.data if self.n else None,
                                           self.hi.ctypes.data if self.hi is not None and self.n else None,
                                           d.ctypes.data if d is not None and self.n else None, self.n_atoms, self.bits_j,
                                           feat.ctypes.data if self.n_atoms else None, so.ctypes.data if so is not None else None,
                                           0 if so is None else so.shape[0] - 1, out.ctypes.data if self.n else None, self.n, int(threads))
        if rc != abi.OK:
            raise ArpeggioCudaError(rc, 'arp_pairs_unpack_packed failed')
        return out


class _BatchInfo:
    """What the engine remembers of a batch upload: the host arrays (kept alive) and the total atom count."""

    def __init__(self, soas, n_atoms):
        self.soas, self.n_atoms = soas, n_atoms
        self.struct_off = True          # a batch: the plane terms do not apply


class ContactEngine:
    def __init__(self, device=0, params=None):
        self._L = lib()
        self._ctx = C.c_void_p()
        rc = self._L.arp_create(int(device), C.byref(self._ctx))
        if rc != abi.OK:
            msg = self._L.arp_last_error(None)
            raise ArpeggioCudaError(rc, (msg or b'arp_create failed').decode())
        self.device = int(device)
        self._soa = None
        self._planes = None
        self.params = None
        self.set_params(params if params is not None else arp_params.make_params())

    # ------------------------------------------------------------------
    def _check(self, rc):
        if rc != abi.OK:
            msg = self._L.arp_last_error(self._ctx)
            raise ArpeggioCudaError(rc, (msg or b'?').decode())

    def close(self):
        if self._ctx:
            self._L.arp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def set_params(self, p):
        """p: abi.ArpParams from params.make_params (thresholds of config.CONTACT_TYPES)."""
        self._check(self._L.arp_set_params(self._ctx, C.byref(p)))
        self.params = p

    def upload_atoms(self, soa, check_finite=True):
        """soa: AtomSoA.  Coordinates must be finite (the reference would compare NaNs false)."""
        if check_finite and soa.n_atoms and not np.isfinite(soa.xyz).all():
            raise ValueError('non-finite atom coordinates')
        a = soa.as_ctypes()
        self._check(self._L.arp_upload_atoms(self._ctx, C.byref(a)))
        self._soa = soa            # keeps the host arrays alive while the copies are in flight

    def upload_atoms_batch(self, soas, check_finite=True):
        """Several independent structures (AtomSoA each, indices local to the structure) as one batch: every structure
        travels with its own DMA and the device concatenates them (arp_upload_atoms_batch), so that run_pairs works on
        all of them in one launch sequence.  Returns the atom offsets [len(soas) + 1] of the structures in the batch:
        the atom indices of the results are global."""
        soas = list(soas)
        if check_finite:
            for soa in soas:
                if soa.n_atoms and not np.isfinite(soa.xyz).all():
                    raise ValueError('non-finite atom coordinates')
        structs = [soa.as_ctypes() for soa in soas]
        arr = (C.POINTER(abi.ArpAtoms) * len(structs))(*[C.pointer(a) for a in structs])
        self._check(self._L.arp_upload_atoms_batch(self._ctx, arr, len(structs)))
        off = np.concatenate([[0], np.cumsum([soa.n_atoms for soa in soas])]).astype(np.int64)
        self._soa = _BatchInfo(soas, int(off[-1]))            # keeps the host arrays alive while the copies are in flight
        return off

    def max_struct_atoms(self):
        """Atoms of the largest structure of the upload: the packed view needs a fifth byte per record beyond 131072."""
        soa = self._soa
        if isinstance(soa, _BatchInfo):
            return max((s.n_atoms for s in soa.soas), default=0)
        if getattr(soa, 'struct_off', None) is not None:
            return int(np.diff(soa.struct_off).max()) if soa.struct_off.shape[0] > 1 else 0
        return soa.n_atoms

    def run_pairs(self):
        """Grid build + pair kernel on the uploaded atoms; returns the number of contact records."""
        n = C.c_uint64()
        self._check(self._L.arp_pairs_run(self._ctx, C.byref(n)))
        return int(n.value)

    def fetch_pairs(self, n, sorted=True, out=None):
        """Record stream of the last run as abi.PAIR_DTYPE; sorted: (i, j) ascending."""
        if out is None:
            out = np.empty(n, dtype=abi.PAIR_DTYPE)
        elif out.dtype != abi.PAIR_DTYPE or out.shape[0] < n or not out.flags['C_CONTIGUOUS']:
            raise ValueError('out must be a C-contiguous PAIR_DTYPE array of at least n records')
        self._check(self._L.arp_pairs_fetch(self._ctx, out.ctypes.data if out.shape[0] else None, out.shape[0],
                                            1 if sorted else 0))
        return out[:n]

    def run_pairs_async(self):
        """Enqueues the job of run_pairs without waiting for it: the next pair_count / fetch_* call does."""
        self._check(self._L.arp_pairs_run_async(self._ctx))

    def pair_count(self):
        n = C.c_uint64()
        self._check(self._L.arp_pairs_count(self._ctx, C.byref(n)))
        return int(n.value)

    def fetch_pairs_compact(self, with_dist=False, out=None, grow=None):
        """The (i, j)-sorted stream of the last run in its compact form (arp_pairs_fetch_compact): returns
        CompactPairs(row_off uint32[N + 1], rec PAIR_C_DTYPE[n], dist float32[n] or None).  8 bytes per record + 4 per
        atom cross PCIe instead of 16 per record; distances on demand (fetch_pairs_dist).  out: a CompactPairs whose
        arrays are reused when they are large enough (pinned buffers of a batch slot); grow(n): called for a larger
        CompactPairs when the stream does not fit `out`."""
        n_atoms = self._soa.n_atoms
        n = C.c_uint64()
        if out is None or out.row_off.shape[0] < n_atoms + 1:
            out = CompactPairs(np.empty(n_atoms + 1, np.uint32), np.empty(0, abi.PAIR_C_DTYPE), None)
        for attempt in range(2):
            dist = out.dist_buf if with_dist else None
            cap = out.rec_buf.shape[0] if dist is None else min(out.rec_buf.shape[0], dist.shape[0])
            rc = self._L.arp_pairs_fetch_compact(self._ctx, out.row_off.ctypes.data, out.rec_buf.ctypes.data if cap else None, cap,
                                                 dist.ctypes.data if dist is not None and cap else None, C.byref(n))
            if rc == abi.E_CAPACITY and attempt == 0:      # the count is known now: size the buffers and fetch again
                m = int(n.value)
                out = grow(m) if grow is not None else CompactPairs(out.row_off, np.empty(m, abi.PAIR_C_DTYPE),
                                                                    np.empty(m, np.float32) if with_dist else None)
                continue
            self._check(rc)
            break
        return out.view(n_atoms, int(n.value), with_dist)

    def fetch_pairs_packed(self, with_dist=False, out=None, grow=None):
        """The (i, j)-sorted stream of the last run in its packed form (arp_pairs_fetch_packed): PackedPairs, 4 bytes per
        record (5 beyond 131072 atoms) + 4 per atom.  out / grow as for fetch_pairs_compact."""
        n_atoms = self._soa.n_atoms
        n, bits, faults = C.c_uint64(), C.c_int32(), C.c_uint32()
        wide = self.max_struct_atoms() > (1 << 17)
        if out is None or out.row_off.shape[0] < n_atoms + 1:
            out = PackedPairs(np.empty(n_atoms + 1, np.uint32), np.empty(0, np.uint32), np.empty(0, np.uint8) if wide else None, None)
        for attempt in range(2):
            dist = out.dist_buf if with_dist else None
            cap = out.lo_buf.shape[0]
            if wide:
                cap = min(cap, out.hi_buf.shape[0]) if out.hi_buf is not None else 0
            if dist is not None:
                cap = min(cap, dist.shape[0])
            rc = self._L.arp_pairs_fetch_packed(self._ctx, out.row_off.ctypes.data, out.lo_buf.ctypes.data if cap else None,
                                                out.hi_buf.ctypes.data if wide and cap else None, cap,
                                                dist.ctypes.data if dist is not None and cap else None, C.byref(n), C.byref(bits), C.byref(faults))
            if rc == abi.E_CAPACITY and attempt == 0:
                m = int(n.value)
                out = grow(m) if grow is not None else PackedPairs(out.row_off, np.empty(m, np.uint32), np.empty(m, np.uint8) if wide else None,
                                                                   np.empty(m, np.float32) if with_dist else None)
                continue
            self._check(rc)
            break
        return out.view(n_atoms, int(n.value), int(bits.value), int(faults.value), with_dist)

    def fetch_pairs_packed_async(self, out, expect=0, with_dist=False):
        """Enqueues the sorted packed view and its copies into `out` (PackedPairs over pinned memory) behind a run that has
        not been waited for (run_pairs_async) and returns at once; fetch_pairs_packed_wait is the step's one wait.  expect:
        the caller's guess of the record count (that many words are copied blindly; the wait fetches what is missing)."""
        n_atoms = self._soa.n_atoms
        wide = self.max_struct_atoms() > (1 << 17)
        if out.row_off.shape[0] < n_atoms + 2:
            raise ValueError('out.row_off must hold n_atoms + 2 entries (the last one is scratch)')
        dist = out.dist_buf if with_dist else None
        cap = out.lo_buf.shape[0]
        if wide:
            cap = min(cap, out.hi_buf.shape[0]) if out.hi_buf is not None else 0
        if dist is not None:
            cap = min(cap, dist.shape[0])
        self._check(self._L.arp_pairs_fetch_packed_async(self._ctx, out.row_off.ctypes.data, out.lo_buf.ctypes.data if cap else None,
                                                         out.hi_buf.ctypes.data if wide and cap else None, cap,
                                                         dist.ctypes.data if dist is not None and cap else None, int(expect)))
        self._pk = (out, with_dist)

    def fetch_pairs_packed_wait(self, grow=None):
        """The wait of fetch_pairs_packed_async: PackedPairs view of its `out`.  A stream that did not fit is fetched again
        into grow(n) (or a fresh pageable buffer) -- the views are still on the device, nothing is recomputed."""
        out, with_dist = self._pk
        self._pk = None
        n, bits, faults = C.c_uint64(), C.c_int32(), C.c_uint32()
        rc = self._L.arp_pairs_fetch_packed_wait(self._ctx, C.byref(n), C.byref(bits), C.byref(faults))
        if rc == abi.E_CAPACITY:
            return self.fetch_pairs_packed(with_dist, out=grow(int(n.value)) if grow is not None else None)
        self._check(rc)
        return out.view(self._soa.n_atoms, int(n.value), int(bits.value), int(faults.value), with_dist)

    def fetch_pairs_dist(self, n, out=None):
        """float32 distances of the sorted stream, same order as the compact records."""
        if out is None:
            out = np.empty(n, np.float32)
        self._check(self._L.arp_pairs_fetch_dist(self._ctx, out.ctypes.data if n else None, out.shape[0]))
        return out[:n]

    def pairs(self, soa=None, sorted=True):
        """upload (optional) + run + fetch: the replacement of the loop at interactions.py:707-936."""
        if soa is not None:
            self.upload_atoms(soa)
        n = self.run_pairs()
        return self.fetch_pairs(n, sorted=sorted)

    def pairs_device_ptr(self):
        p = C.c_void_p()
        self._check(self._L.arp_pairs_device_ptr(self._ctx, C.byref(p)))
        return p.value

    # ------------------------------------------------------------------
    def upload_planes(self, rings, amides):
        """rings: PlaneSoA float64; amides: PlaneSoA float32 (either may be None / empty)."""
        r = rings.as_ctypes() if rings is not None else None
        a = amides.as_ctypes() if amides is not None else None
        self._check(self._L.arp_upload_planes(self._ctx, C.byref(r) if r is not None else None,
                                              C.byref(a) if a is not None else None))
        self._planes = (rings, amides)

    def _plane_term(self, run, fetch, dtype):
        n = C.c_uint64()
        self._check(run(self._ctx, C.byref(n)))
        out = np.empty(int(n.value), dtype=dtype)
        self._check(fetch(self._ctx, out.ctypes.data if out.shape[0] else None, out.shape[0]))
        return out

    def planes_all(self):
        """All four plane terms in one launch sequence (arp_planes_run_all): dict of record arrays
        ring_ring / atom_ring / amide_amide / amide_ring (atom_ring is None when no single structure is uploaded)."""
        n = (C.c_uint64 * 4)()
        self._check(self._L.arp_planes_run_all(self._ctx, n))
        have_atoms = self._soa is not None and getattr(self._soa, 'struct_off', None) is None
        out = {}
        for k, (name, fetch, dtype) in enumerate((('ring_ring', self._L.arp_ring_ring_fetch, abi.PLANE_PAIR_DTYPE),
                                                  ('atom_ring', self._L.arp_atom_ring_fetch, abi.ATOM_PLANE_DTYPE),
                                                  ('amide_amide', self._L.arp_amide_amide_fetch, abi.PLANE_PAIR_DTYPE),
                                                  ('amide_ring', self._L.arp_amide_ring_fetch, abi.PLANE_PAIR_DTYPE))):
            if name == 'atom_ring' and not have_atoms:
                out[name] = None
                continue
            a = np.empty(int(n[k]), dtype=dtype)
            self._check(fetch(self._ctx, a.ctypes.data if a.shape[0] else None, a.shape[0]))
            out[name] = a
        return out

    def ring_ring(self):
        return self._plane_term(self._L.arp_ring_ring_run, self._L.arp_ring_ring_fetch, abi.PLANE_PAIR_DTYPE)

    def atom_ring(self):
        return self._plane_term(self._L.arp_atom_ring_run, self._L.arp_atom_ring_fetch, abi.ATOM_PLANE_DTYPE)

    def amide_amide(self):
        return self._plane_term(self._L.arp_amide_amide_run, self._L.arp_amide_amide_fetch, abi.PLANE_PAIR_DTYPE)

    def amide_ring(self):
        return self._plane_term(self._L.arp_amide_ring_run, self._L.arp_amide_ring_fetch, abi.PLANE_PAIR_DTYPE)

    # ------------------------------------------------------------------
    def atom_sifts(self):
        """arp_atom_sift[N] of the last run_pairs: the per-atom SIFt words, integer SIFts and hbond / polar
        counters the reference's pair loop leaves on the atoms (utils.py:182-242, interactions.py:822-852)."""
        n = self._soa.n_atoms
        out = np.zeros(n, dtype=abi.ATOM_SIFT_DTYPE)
        self._check(self._L.arp_atom_sifts_run(self._ctx))
        self._check(self._L.arp_atom_sifts_fetch(self._ctx, out.ctypes.data if n else None, n))
        return out

    def ring_nearest_atom(self, xyz, centers, radius=3.0):
        """(atom index or -1, float64 distance) of the atom closest to every ring centroid within `radius`
        (_assign_aromatic_rings_to_residues, interactions.py:1453-1492).  xyz: float32[N][3] of the atoms
        searched (the reference searches s_atoms), centers: float64[R][3]."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
        r = centers.shape[0]
        atom = np.full(r, -1, dtype=np.int32)
        dist = np.zeros(r, dtype=np.float64)
        if r:
            self._check(self._L.arp_ring_nearest_atom(self._ctx, xyz.ctypes.data if xyz.shape[0] else None, xyz.shape[0],
                                                      centers.ctypes.data, r, float(radius), atom.ctypes.data, dist.ctypes.data))
        return atom, dist

    def flag_within(self, radius):
        """uint8[N]: atom selected or within `radius` of a selected atom (interactions.py:1420-1424)."""
        n = self._soa.n_atoms
        out = np.zeros(n, dtype=np.uint8)
        self._check(self._L.arp_flag_within(self._ctx, float(radius), out.ctypes.data if n else None, n))
        return out

    def sync(self):
        self._check(self._L.arp_sync(self._ctx))

    def stats(self):
        s = abi.ArpStats()
        self._check(self._L.arp_get_stats(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in abi.ArpStats._fields_}

    def launch_count(self):
        return int(self._L.arp_launch_count(self._ctx))

    def memcpy_probe(self, h2d_bytes, d2h_bytes, iters=20):
        """GB/s of pinned cudaMemcpyAsync host -> device and device -> host, both directions at once (bench hook)."""
        ms = (C.c_float * 2)()
        self._check(self._L.arp_memcpy_probe(self._ctx, int(h2d_bytes), int(d2h_bytes), int(iters), ms))
        return {'h2d_gbs': h2d_bytes * iters / (ms[0] * 1e6) if ms[0] > 0 else 0.0,
                'd2h_gbs': d2h_bytes * iters / (ms[1] * 1e6) if ms[1] > 0 else 0.0}

    def time_pairs(self, iters, flush_l2=True):
        """Mean CUDA-event time (ms) of the whole atom-atom job over `iters` runs on the resident inputs."""
        ms = C.c_float()
        self._check(self._L.arp_timing_iters(self._ctx, int(iters), 1 if flush_l2 else 0, C.byref(ms)))
        return float(ms.value)
