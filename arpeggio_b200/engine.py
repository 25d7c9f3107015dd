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
    from .soa import AtomSoA
    names = ('xyz', 'feat', 'res_id', 'rad_class', 'vdw', 'cov', 'res_prev', 'res_next', 'res_flags',
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
    return AtomSoA(**fields, _keep=[block])


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

    def time_pairs(self, iters, flush_l2=True):
        """Mean CUDA-event time (ms) of the whole atom-atom job over `iters` runs on the resident inputs."""
        ms = C.c_float()
        self._check(self._L.arp_timing_iters(self._ctx, int(iters), 1 if flush_l2 else 0, C.byref(ms)))
        return float(ms.value)
