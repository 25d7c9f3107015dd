"""ctypes mirror of ``include/arpeggio_cuda.h``.

Only layouts and constants live here (no library is loaded), so that the product
bindings (:mod:`arpeggio_b200._lib`) and the test-only oracle wrapper
(``oracle/oracle.py``) describe the boundary with one set of definitions.
Field order and types must match the header exactly; ``tests/test_abi.py``
checks sizes and offsets against a C probe compiled from the header.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 6

# ---- error codes ---------------------------------------------------------
OK = 0
E_INVALID_ARG = -1
E_CUDA = -2
E_OOM = -3
E_CAPACITY = -4
E_NOT_READY = -5
E_NO_DEVICE = -6

# ---- SIFt bits: order of `contacts` in InteractionComplex.get_contacts
# (arpeggio/core/interactions.py:178-180)
SIFT_NAMES = ('clash', 'covalent', 'vdw_clash', 'vdw', 'proximal', 'hbond', 'weak_hbond',
              'xbond', 'ionic', 'metal_complex', 'aromatic', 'hydrophobic', 'carbonyl',
              'polar', 'weak_polar')
SIFT_NBITS = 15
CLASS_SHIFT = 16
CLASS_MASK = 0x7
# interacting_entities strings (interactions.py:669-685; planes :988-997)
CLASS_NAMES = ('INTRA_NON_SELECTION', 'INTRA_SELECTION', 'INTER', 'SELECTION_WATER',
               'NON_SELECTION_WATER', 'WATER_WATER', 'INTRA_BINDING_SITE')
PAIR_FAULT_XBOND_NO_NBR = 1 << 31   # utils.py:173 would dereference None

# ---- atom feature bits: the 12 keys of config.ATOM_TYPES (config.py:53-145) in bit order
ATOM_TYPE_KEYS = ('hbond acceptor', 'hbond donor', 'weak hbond acceptor', 'weak hbond donor',
                  'xbond acceptor', 'xbond donor', 'pos ionisable', 'neg ionisable',
                  'hydrophobe', 'carbonyl oxygen', 'carbonyl carbon', 'aromatic')
F_HBOND_ACCEPTOR = 1 << 0
F_HBOND_DONOR = 1 << 1
F_WEAK_HBOND_ACCEPTOR = 1 << 2
F_WEAK_HBOND_DONOR = 1 << 3
F_XBOND_ACCEPTOR = 1 << 4
F_XBOND_DONOR = 1 << 5
F_POS_IONISABLE = 1 << 6
F_NEG_IONISABLE = 1 << 7
F_HYDROPHOBE = 1 << 8
F_CARBONYL_OXYGEN = 1 << 9
F_CARBONYL_CARBON = 1 << 10
F_AROMATIC = 1 << 11
F_IS_METAL = 1 << 12
F_IS_HALOGEN = 1 << 13
F_IS_WATER = 1 << 14
F_IN_SELECTION = 1 << 15
F_ELEM_H = 1 << 16
F_ELEM_C = 1 << 17
F_MET_SULPHUR = 1 << 18
F_HAS_XNBR = 1 << 19

R_IS_POLYPEPTIDE = 1 << 0
R_HAS_LINKS = 1 << 1

P_IN_SELECTION = 1 << 0
P_IN_SELECTION_PLUS = 1 << 1

# plane-plane geometry labels (interactions.py:1129-1148); index 9 is the empty string the
# reference produces when an angle is NaN
GEOM_NAMES = ('FF', 'OF', 'EE', 'FT', 'OT', 'ET', 'FE', 'OE', 'EF', '')
G_NONE = 9
# atom-plane labels (interactions.py:1007-1024)
AP_NAMES = ('CARBONPI', 'CATIONPI', 'DONORPI', 'HALOGENPI', 'METSULPHURPI')


class ArpParams(C.Structure):
    _fields_ = [
        ('interacting_cutoff', C.c_double),
        ('vdw_comp', C.c_double),
        ('include_sequence_adjacent', C.c_int32),
        ('blas_fma', C.c_int32),
        ('h_vdw', C.c_double),
        ('dist_max', C.c_double),
        ('hbond_polar_dist', C.c_double),
        ('weak_polar_dist', C.c_double),
        ('ionic_dist', C.c_double),
        ('carbonyl_dist', C.c_double),
        ('aromatic_dist', C.c_double),
        ('hydrophobic_dist', C.c_double),
        ('metal_dist', C.c_double),
        ('hbond_angle', C.c_double),
        ('weak_hbond_angle', C.c_double),
        ('cx_angle_min', C.c_double),
        ('cx_angle_max', C.c_double),
        ('xbond_angle', C.c_double),
        ('ring_centroid_dist', C.c_double),
        ('atom_ring_dist', C.c_double),
        ('met_sulphur_dist', C.c_double),
        ('amide_centroid_dist', C.c_double),
        ('plane_bins_deg', C.c_double * 3),
        ('cos_hbond', C.c_double),
        ('cos_weak_hbond', C.c_double),
        ('cos_cx_min', C.c_double),
        ('cos_cx_max', C.c_double),
        ('cos_xbond_f32', C.c_float),
        ('_pad0', C.c_float),
        ('cos_split_f64', C.c_double),
        ('cos_pos_f64', C.c_double * 3),
        ('cos_neg_f64', C.c_double * 3),
        ('cos_split_f32', C.c_float),
        ('cos_pos_f32', C.c_float * 3),
        ('cos_neg_f32', C.c_float * 3),
        ('_pad1', C.c_float),
    ]


class ArpAtoms(C.Structure):
    _fields_ = [
        ('n_atoms', C.c_int32),
        ('n_residues', C.c_int32),
        ('n_rad_classes', C.c_int32),
        ('n_structures', C.c_int32),
        ('xyz', C.c_void_p),
        ('feat', C.c_void_p),
        ('res_id', C.c_void_p),
        ('rad_class', C.c_void_p),
        ('vdw', C.c_void_p),
        ('cov', C.c_void_p),
        ('res_prev', C.c_void_p),
        ('res_next', C.c_void_p),
        ('res_flags', C.c_void_p),
        ('bond_off', C.c_void_p),
        ('bond_nbr', C.c_void_p),
        ('h_off', C.c_void_p),
        ('h_xyz', C.c_void_p),
        ('xnbr_xyz', C.c_void_p),
        ('struct_off', C.c_void_p),
        ('bond_cnt', C.c_void_p),      # wire forms (optional)
        ('h_cnt', C.c_void_p),
        ('h_fix', C.c_void_p),
        ('xnbr_idx', C.c_void_p),
        ('h_fix_scale', C.c_double),
        ('n_bond_nbr', C.c_int32),
        ('n_h', C.c_int32),
        ('n_xnbr', C.c_int32),
        ('_pad', C.c_int32),
    ]


class ArpPlanes(C.Structure):
    _fields_ = [
        ('n', C.c_int32),
        ('is_f32', C.c_int32),
        ('center', C.c_void_p),
        ('normal', C.c_void_p),
        ('res_id', C.c_void_p),
        ('flags', C.c_void_p),
    ]


class ArpStats(C.Structure):
    _fields_ = [
        ('n_pairs', C.c_uint64),
        ('n_candidates', C.c_uint64),
        ('n_cells', C.c_uint64),
        ('n_cells_nonempty', C.c_uint64),
        ('input_bytes', C.c_uint64),
        ('output_bytes', C.c_uint64),
        ('ms_total', C.c_float),
        ('ms_grid', C.c_float),
        ('ms_search', C.c_float),
        ('ms_classify', C.c_float),
        ('ms_hscan', C.c_float),
        ('ms_pairs', C.c_float),
        ('faults', C.c_uint32),
        ('_pad', C.c_uint32),
    ]


FAULT_NONFINITE = 1
FAULT_HANDOFF = 2


# record layouts as NumPy structured dtypes (arp_pair / arp_plane_pair / arp_atom_plane)
PAIR_DTYPE = np.dtype([('i', '<i4'), ('j', '<i4'), ('mask', '<u4'), ('dist', '<f4')])
PAIR_C_DTYPE = np.dtype([('j', '<i4'), ('mask', '<u4')])      # arp_pair_c: compact view of the sorted stream
PLANE_PAIR_DTYPE = np.dtype([('a', '<i4'), ('b', '<i4'), ('code', '<u4'), ('_pad', '<u4'), ('dist', '<f8')])
ATOM_PLANE_DTYPE = np.dtype([('atom', '<i4'), ('ring', '<i4'), ('code', '<u4'), ('_pad', '<u4'), ('dist', '<f8')])
ATOM_SIFT_DTYPE = np.dtype([('sift', '<u2', (4,)), ('integer_sift', '<u4', (4,)), ('hbonds', '<u4', (4,)), ('polars', '<u4', (4,))])
assert PAIR_DTYPE.itemsize == 16 and PLANE_PAIR_DTYPE.itemsize == 24 and ATOM_PLANE_DTYPE.itemsize == 24
assert ATOM_SIFT_DTYPE.itemsize == 56
SIFT_CATEGORIES = ('', '_inter_only', '_intra_only', '_water_only')     # attribute suffixes of the four categories

# every symbol include/arpeggio_cuda.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = (
    'arp_abi_version', 'arp_device_count', 'arp_create', 'arp_destroy', 'arp_last_error',
    'arp_params_default', 'arp_set_params', 'arp_host_alloc', 'arp_host_free',
    'arp_upload_atoms', 'arp_upload_atoms_batch', 'arp_pairs_run', 'arp_pairs_fetch', 'arp_pairs_device_ptr',
    'arp_pairs_run_async', 'arp_pairs_count', 'arp_pairs_fetch_compact', 'arp_pairs_fetch_dist', 'arp_pairs_fetch_packed', 'arp_pairs_unpack_packed', 'arp_pairs_fetch_packed_async', 'arp_pairs_fetch_packed_wait', 'arp_pairs_unpack',
    'arp_upload_planes', 'arp_ring_ring_run', 'arp_ring_ring_fetch', 'arp_atom_ring_run',
    'arp_atom_ring_fetch', 'arp_amide_amide_run', 'arp_amide_amide_fetch', 'arp_amide_ring_run',
    'arp_amide_ring_fetch', 'arp_planes_run_all', 'arp_atom_sifts_run', 'arp_atom_sifts_fetch', 'arp_ring_nearest_atom', 'arp_pairs_json_size', 'arp_pairs_json_write', 'arp_flag_within', 'arp_sync', 'arp_get_stats', 'arp_timing_iters', 'arp_launch_count', 'arp_memcpy_probe',
)


def ptr(a):
    """Address of a C-contiguous NumPy array (or None -> NULL)."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data
