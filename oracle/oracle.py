"""ctypes front-end of the CPU oracle (``oracle/arp_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``arpeggio_b200/`` may import this module (tests/test_no_oracle_in_product.py checks).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from arpeggio_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'liborc.so')
    src = os.path.join(_HERE, 'arp_oracle.c')
    hdr = os.path.join(_HERE, '..', 'include', 'arpeggio_cuda.h')
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(['make', '-C', _HERE, '-B', 'liborc.so'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
        L.orc_pairs.argtypes = [C.POINTER(abi.ArpAtoms), C.POINTER(abi.ArpParams), C.POINTER(vp), u64p, u64p]
        L.orc_pairs_bruteforce.argtypes = [C.POINTER(abi.ArpAtoms), C.POINTER(abi.ArpParams), C.POINTER(vp), u64p]
        L.orc_classify.argtypes = [C.POINTER(abi.ArpAtoms), C.POINTER(abi.ArpParams), vp, vp, C.c_int64, vp, vp]
        L.orc_flag_within.argtypes = [C.POINTER(abi.ArpAtoms), C.c_double, vp]
        L.orc_ring_ring.argtypes = [C.POINTER(abi.ArpPlanes), C.POINTER(abi.ArpParams), C.POINTER(vp), u64p]
        L.orc_amide_amide.argtypes = [C.POINTER(abi.ArpPlanes), C.POINTER(abi.ArpParams), C.POINTER(vp), u64p]
        L.orc_amide_ring.argtypes = [C.POINTER(abi.ArpPlanes), C.POINTER(abi.ArpPlanes), C.POINTER(abi.ArpParams),
                                     C.POINTER(vp), u64p]
        L.orc_atom_ring.argtypes = [C.POINTER(abi.ArpAtoms), C.POINTER(abi.ArpPlanes), C.POINTER(abi.ArpParams),
                                    C.POINTER(vp), u64p]
        L.orc_atom_sifts.argtypes = [vp, C.c_uint64, C.c_int, vp]
        L.orc_ring_nearest.argtypes = [vp, C.c_int, vp, C.c_int, C.c_double, C.c_int, vp, vp]
        L.orc_free.argtypes = [vp]
        L.orc_free.restype = None
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_dot3_f64.argtypes = [dp, dp, C.c_int]; L.orc_dot3_f64.restype = C.c_double
        L.orc_norm3_f64.argtypes = [dp, C.c_int]; L.orc_norm3_f64.restype = C.c_double
        L.orc_dot3_f32.argtypes = [fp, fp]; L.orc_dot3_f32.restype = C.c_float
        L.orc_norm3_f32.argtypes = [fp]; L.orc_norm3_f32.restype = C.c_float
        L.orc_dist_f32_pub.argtypes = [fp, fp]; L.orc_dist_f32_pub.restype = C.c_float
        L.orc_fold_deg_f64.argtypes = [C.c_double]; L.orc_fold_deg_f64.restype = C.c_double
        L.orc_fold_deg_f32.argtypes = [C.c_float]; L.orc_fold_deg_f32.restype = C.c_float
        L.orc_get_angle_fdf.argtypes = [fp, dp, fp]; L.orc_get_angle_fdf.restype = C.c_double
        L.orc_get_angle_ffd.argtypes = [fp, fp, dp]; L.orc_get_angle_ffd.restype = C.c_double
        L.orc_get_angle_fff.argtypes = [fp, fp, fp]; L.orc_get_angle_fff.restype = C.c_double
        _LIB = L
    return _LIB


def _take(ptr, n, dtype):
    n = int(n)
    if n == 0 or not ptr.value:
        if ptr.value:
            lib().orc_free(ptr)
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=dtype).copy()
    lib().orc_free(ptr)
    return out


def _check(rc):
    if rc != 0:
        raise RuntimeError(f'oracle failed with code {rc}')


def pairs(soa, params, bruteforce=False, with_counts=False):
    """Oracle of upload_atoms + pairs_run + pairs_fetch(sorted): records sorted by (i, j)."""
    a = soa.as_ctypes()
    out, n, nw = C.c_void_p(), C.c_uint64(), C.c_uint64()
    if bruteforce:
        _check(lib().orc_pairs_bruteforce(C.byref(a), C.byref(params), C.byref(out), C.byref(n)))
    else:
        _check(lib().orc_pairs(C.byref(a), C.byref(params), C.byref(out), C.byref(n), C.byref(nw)))
    rec = _take(out, n.value, abi.PAIR_DTYPE)
    if bruteforce:
        rec = rec[np.lexsort((rec['j'], rec['i']))]
    return (rec, int(nw.value)) if with_counts else rec


def classify(soa, params, b, e):
    """The loop body of _calculate_atom_contacts for explicit (bgn, end) index lists."""
    a = soa.as_ctypes()
    b = np.ascontiguousarray(b, np.int32)
    e = np.ascontiguousarray(e, np.int32)
    out = np.zeros(b.shape[0], dtype=abi.PAIR_DTYPE)
    emitted = np.zeros(b.shape[0], dtype=np.uint8)
    _check(lib().orc_classify(C.byref(a), C.byref(params), b.ctypes.data, e.ctypes.data, b.shape[0],
                              out.ctypes.data, emitted.ctypes.data))
    return out, emitted.astype(bool)


def atom_sifts(records, n_atoms):
    """Oracle of atom_sifts(): the pair loop's side effects on the atoms, replayed over `records` in their order."""
    rec = np.ascontiguousarray(records, dtype=abi.PAIR_DTYPE)
    out = np.zeros(int(n_atoms), dtype=abi.ATOM_SIFT_DTYPE)
    _check(lib().orc_atom_sifts(rec.ctypes.data if rec.shape[0] else None, rec.shape[0], int(n_atoms),
                                out.ctypes.data if n_atoms else None))
    return out


def ring_nearest_atom(xyz, centers, radius, params):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
    atom = np.full(centers.shape[0], -1, dtype=np.int32)
    dist = np.zeros(centers.shape[0], dtype=np.float64)
    _check(lib().orc_ring_nearest(xyz.ctypes.data, xyz.shape[0], centers.ctypes.data, centers.shape[0], float(radius),
                                  int(params.blas_fma), atom.ctypes.data, dist.ctypes.data))
    return atom, dist


def flag_within(soa, radius):
    a = soa.as_ctypes()
    flags = np.zeros(soa.n_atoms, dtype=np.uint8)
    _check(lib().orc_flag_within(C.byref(a), float(radius), flags.ctypes.data))
    return flags


def ring_ring(rings, params):
    r = rings.as_ctypes()
    out, n = C.c_void_p(), C.c_uint64()
    _check(lib().orc_ring_ring(C.byref(r), C.byref(params), C.byref(out), C.byref(n)))
    return _take(out, n.value, abi.PLANE_PAIR_DTYPE)


def amide_amide(amides, params):
    r = amides.as_ctypes()
    out, n = C.c_void_p(), C.c_uint64()
    _check(lib().orc_amide_amide(C.byref(r), C.byref(params), C.byref(out), C.byref(n)))
    return _take(out, n.value, abi.PLANE_PAIR_DTYPE)


def amide_ring(amides, rings, params):
    a, r = amides.as_ctypes(), rings.as_ctypes()
    out, n = C.c_void_p(), C.c_uint64()
    _check(lib().orc_amide_ring(C.byref(a), C.byref(r), C.byref(params), C.byref(out), C.byref(n)))
    return _take(out, n.value, abi.PLANE_PAIR_DTYPE)


def atom_ring(soa, rings, params):
    a, r = soa.as_ctypes(), rings.as_ctypes()
    out, n = C.c_void_p(), C.c_uint64()
    _check(lib().orc_atom_ring(C.byref(a), C.byref(r), C.byref(params), C.byref(out), C.byref(n)))
    return _take(out, n.value, abi.ATOM_PLANE_DTYPE)
