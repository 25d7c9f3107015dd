"""Loader and ctypes prototypes of ``libarpeggio_cuda.so`` (include/arpeggio_cuda.h).

There is no CPU fallback: if the library has not been built (``python __graft_entry__.py`` or
``make -C arpeggio_b200/csrc``) or cannot be loaded, importing callers get an ``ImportError``
that says so.
"""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ARPEGGIO_CUDA_LIB') or os.path.join(_HERE, 'libarpeggio_cuda.so')   # override: A/B builds
_LIB = None


class ArpeggioCudaError(RuntimeError):
    """A call into libarpeggio_cuda.so failed (negative ARP_E_* code)."""

    def __init__(self, code, message):
        super().__init__(f'libarpeggio_cuda: {message} [code {code}]')
        self.code = code


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `make -C arpeggio_b200/csrc` '
                          '(nvcc, sm_100a). arpeggio_b200 has no CPU fallback.')
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as err:
        raise ImportError(f'cannot load {LIB_PATH}: {err}. arpeggio_b200 has no CPU fallback.') from err
    vp, u64, u64p, i32 = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int
    proto = {
        'arp_abi_version': (i32, []),
        'arp_device_count': (i32, []),
        'arp_create': (i32, [i32, C.POINTER(vp)]),
        'arp_destroy': (None, [vp]),
        'arp_last_error': (C.c_char_p, [vp]),
        'arp_params_default': (i32, [C.POINTER(abi.ArpParams)]),
        'arp_set_params': (i32, [vp, C.POINTER(abi.ArpParams)]),
        'arp_host_alloc': (i32, [C.POINTER(vp), u64]),
        'arp_host_free': (i32, [vp]),
        'arp_upload_atoms': (i32, [vp, C.POINTER(abi.ArpAtoms)]),
        'arp_upload_atoms_batch': (i32, [vp, C.POINTER(C.POINTER(abi.ArpAtoms)), i32]),
        'arp_pairs_run': (i32, [vp, u64p]),
        'arp_pairs_fetch': (i32, [vp, vp, u64, i32]),
        'arp_pairs_device_ptr': (i32, [vp, C.POINTER(vp)]),
        'arp_pairs_run_async': (i32, [vp]),
        'arp_pairs_count': (i32, [vp, u64p]),
        'arp_pairs_fetch_compact': (i32, [vp, vp, vp, u64, vp, u64p]),
        'arp_pairs_fetch_dist': (i32, [vp, vp, u64]),
        'arp_pairs_fetch_packed': (i32, [vp, vp, vp, vp, u64, vp, u64p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]),
        'arp_pairs_unpack_packed': (i32, [vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, u64, i32]),
        'arp_pairs_fetch_packed_async': (i32, [vp, vp, vp, vp, u64, vp, u64]),
        'arp_pairs_fetch_packed_wait': (i32, [vp, u64p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]),
        'arp_pairs_unpack': (i32, [vp, vp, vp, i32, vp, u64]),
        'arp_upload_planes': (i32, [vp, C.POINTER(abi.ArpPlanes), C.POINTER(abi.ArpPlanes)]),
        'arp_ring_ring_run': (i32, [vp, u64p]),
        'arp_ring_ring_fetch': (i32, [vp, vp, u64]),
        'arp_atom_ring_run': (i32, [vp, u64p]),
        'arp_atom_ring_fetch': (i32, [vp, vp, u64]),
        'arp_amide_amide_run': (i32, [vp, u64p]),
        'arp_amide_amide_fetch': (i32, [vp, vp, u64]),
        'arp_amide_ring_run': (i32, [vp, u64p]),
        'arp_amide_ring_fetch': (i32, [vp, vp, u64]),
        'arp_planes_run_all': (i32, [vp, u64p]),
        'arp_ring_nearest_atom': (i32, [vp, vp, i32, vp, i32, C.c_double, vp, vp]),
        'arp_atom_sifts_run': (i32, [vp]),
        'arp_atom_sifts_fetch': (i32, [vp, vp, u64]),
        'arp_pairs_json_size': (i32, [vp, u64, i32, vp, i32, u64p]),
        'arp_pairs_json_write': (i32, [vp, u64, i32, vp, vp, i32, vp, u64, u64p]),
        'arp_flag_within': (i32, [vp, C.c_double, vp, u64]),
        'arp_sync': (i32, [vp]),
        'arp_get_stats': (i32, [vp, C.POINTER(abi.ArpStats)]),
        'arp_timing_iters': (i32, [vp, i32, i32, C.POINTER(C.c_float)]),
        'arp_launch_count': (u64, [vp]),
        'arp_memcpy_probe': (i32, [vp, u64, u64, i32, C.POINTER(C.c_float)]),
    }
    assert set(proto) == set(abi.EXPORTED_SYMBOLS)
    for name, (res, args) in proto.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    got = L.arp_abi_version()
    if got != abi.ABI_VERSION:
        raise ImportError(f'{LIB_PATH} has ABI version {got}, the Python side expects {abi.ABI_VERSION}: rebuild it')
    _LIB = L
    return L
