"""The C ABI boundary without a GPU: the ctypes mirror matches include/arpeggio_cuda.h, the built
library exports every declared symbol, and the product fails loudly when no device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from arpeggio_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'arpeggio_cuda.h')


def _probe(tmp_path, structs):
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, cls in structs:
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['return 0;}']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.check_call(['gcc', '-std=c11', '-o', str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True)
    return {tuple(l.split()[:2]): int(l.split()[2]) for l in out.splitlines()}


def test_struct_layouts_match_header(tmp_path):
    structs = [('arp_params', abi.ArpParams), ('arp_atoms', abi.ArpAtoms), ('arp_planes', abi.ArpPlanes),
               ('arp_stats', abi.ArpStats)]
    got = _probe(tmp_path, structs)
    for cname, cls in structs:
        assert got[(cname, 'size')] == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert got[(cname, f)] == getattr(cls, f).offset, (cname, f)


def test_record_layouts_match_header(tmp_path):
    class Pair(C.Structure):
        _fields_ = [('i', C.c_int32), ('j', C.c_int32), ('mask', C.c_uint32), ('dist', C.c_float)]

    class PlanePair(C.Structure):
        _fields_ = [('a', C.c_int32), ('b', C.c_int32), ('code', C.c_uint32), ('_pad', C.c_uint32), ('dist', C.c_double)]

    class AtomPlane(C.Structure):
        _fields_ = [('atom', C.c_int32), ('ring', C.c_int32), ('code', C.c_uint32), ('_pad', C.c_uint32), ('dist', C.c_double)]

    class PairC(C.Structure):
        _fields_ = [('j', C.c_int32), ('mask', C.c_uint32)]

    got = _probe(tmp_path, [('arp_pair', Pair), ('arp_plane_pair', PlanePair), ('arp_atom_plane', AtomPlane),
                            ('arp_pair_c', PairC)])
    for cname, cls, dt in (('arp_pair', Pair, abi.PAIR_DTYPE), ('arp_plane_pair', PlanePair, abi.PLANE_PAIR_DTYPE),
                           ('arp_atom_plane', AtomPlane, abi.ATOM_PLANE_DTYPE), ('arp_pair_c', PairC, abi.PAIR_C_DTYPE)):
        assert got[(cname, 'size')] == dt.itemsize
        for f, _ in cls._fields_:
            assert got[(cname, f)] == dt.fields[f][1], (cname, f)


def test_constants_match_header():
    text = open(HEADER).read()
    defs = dict(re.findall(r'#define\s+(ARP_\w+)\s+\(?(-?\w+)', text))
    assert int(defs['ARP_ABI_VERSION']) == abi.ABI_VERSION
    for k, name in enumerate(abi.SIFT_NAMES):
        key = {'metal_complex': 'METAL'}.get(name, name.upper())
        assert int(defs['ARP_SIFT_' + key]) == k
    for name in ('OK', 'E_INVALID_ARG', 'E_CUDA', 'E_OOM', 'E_CAPACITY', 'E_NOT_READY', 'E_NO_DEVICE'):
        assert int(defs['ARP_' + name]) == getattr(abi, name)
    bits = dict(re.findall(r'#define\s+(ARP_[FRP]_\w+)\s+\(1u << (\d+)\)', text))
    for name, sh in bits.items():
        assert getattr(abi, name[4:]) == 1 << int(sh), name
    for k, name in enumerate(abi.CLASS_NAMES):
        assert int(defs['ARP_CLASS_' + name]) == k


def test_header_declares_what_python_binds():
    text = open(HEADER).read()
    declared = set(re.findall(r'\b(arp_[a-z_0-9]+)\s*\(', text))
    assert declared == set(abi.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    from arpeggio_b200 import _lib
    L = _lib.lib()
    for name in abi.EXPORTED_SYMBOLS:
        assert hasattr(L, name), name
    assert L.arp_abi_version() == abi.ABI_VERSION


def test_no_device_means_error_not_fallback():
    from arpeggio_b200 import _lib
    from arpeggio_b200.engine import ContactEngine
    L = _lib.lib()
    if L.arp_device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(_lib.ArpeggioCudaError) as ei:
        ContactEngine(0)
    assert ei.value.code == abi.E_NO_DEVICE
    assert 'no CPU fallback' in str(ei.value)


def test_default_params_from_c_match_python_thresholds():
    """arp_params_default (C library acos) and params.make_params (host NumPy arccos) hold the same
    thresholds; the cosine images may differ in the last ulps (different arccos implementations)."""
    from arpeggio_b200 import _lib, params
    p = abi.ArpParams()
    assert _lib.lib().arp_params_default(C.byref(p)) == 0
    q = params.make_params()
    for f, _ in abi.ArpParams._fields_:
        a, b = getattr(p, f), getattr(q, f)
        if f.startswith('cos_'):
            a = np.array(list(a) if hasattr(a, '__len__') else [a], dtype=np.float64)
            b = np.array(list(b) if hasattr(b, '__len__') else [b], dtype=np.float64)
            assert np.allclose(a, b, rtol=0, atol=3e-7), f
        elif hasattr(a, '__len__'):
            assert list(a) == list(b), f
        else:
            assert a == b, f


def test_unpack_compact_records_on_the_host():
    """arp_pairs_unpack needs no device: rows -> (i, j, mask, dist) records; malformed offsets are refused."""
    from arpeggio_b200.engine import CompactPairs
    from arpeggio_b200 import _lib
    rng = np.random.default_rng(3)
    n_atoms = 57
    per_row = rng.integers(0, 6, size=n_atoms)
    per_row[[0, 13, n_atoms - 1]] = 0                       # empty rows at the ends and in the middle
    row_off = np.concatenate([[0], np.cumsum(per_row)]).astype(np.uint32)
    n = int(row_off[-1])
    rec = np.zeros(n, abi.PAIR_C_DTYPE)
    rec['j'] = rng.integers(0, n_atoms, size=n)
    rec['mask'] = rng.integers(0, 1 << 19, size=n)
    dist = rng.random(n).astype(np.float32)
    cp = CompactPairs(row_off, rec, dist).view(n_atoms, n, True)
    out = cp.to_records()
    assert np.array_equal(out['i'], np.repeat(np.arange(n_atoms), per_row))
    assert np.array_equal(out['j'], rec['j']) and np.array_equal(out['mask'], rec['mask'])
    assert np.array_equal(out['dist'].view(np.uint32), dist.view(np.uint32))
    assert np.all(CompactPairs(row_off, rec, None).view(n_atoms, n, False).to_records()['dist'] == 0)
    assert cp.nbytes == 4 * (n_atoms + 1) + 12 * n
    L = _lib.lib()
    small = np.empty(max(n - 1, 0), abi.PAIR_DTYPE)
    assert L.arp_pairs_unpack(row_off.ctypes.data, rec.ctypes.data, None, n_atoms, small.ctypes.data, small.shape[0]) == abi.E_CAPACITY
    bad = row_off.copy()
    bad[5], bad[6] = bad[6] + 1, bad[5]                     # descending offsets
    big = np.empty(n + 8, abi.PAIR_DTYPE)
    assert L.arp_pairs_unpack(bad.ctypes.data, rec.ctypes.data, None, n_atoms, big.ctypes.data, big.shape[0]) == abi.E_INVALID_ARG
    assert L.arp_pairs_unpack(None, None, None, 0, None, 0) == abi.OK


def test_unpack_packed_records_on_the_host():
    """arp_pairs_unpack_packed needs no device: word -> (i, j, mask | class, dist), class from the feat words exactly as
    __get_contact_type's six ifs give it; narrow (32-bit) and wide (40-bit) words."""
    from arpeggio_b200.engine import PackedPairs
    rng = np.random.default_rng(4)
    for n_atoms, bits in ((1000, 10), (200_000, 18)):
        per_row = rng.integers(0, 4, size=n_atoms)
        row_off = np.concatenate([[0], np.cumsum(per_row)]).astype(np.uint32)
        n = int(row_off[-1])
        i = np.repeat(np.arange(n_atoms), per_row)
        j = rng.integers(0, n_atoms, size=n).astype(np.int64)
        mask = rng.integers(0, 1 << 15, size=n).astype(np.int64)
        word = j | (mask << bits)
        feat = (rng.integers(0, 4, size=n_atoms).astype(np.uint32) << 14)          # water (bit 14) and selection (bit 15) bits
        dist = rng.random(n).astype(np.float32)
        pk = PackedPairs(row_off, (word & 0xffffffff).astype(np.uint32), (word >> 32).astype(np.uint8) if bits + 15 > 32 else None,
                         dist).view(n_atoms, n, bits, 0, True)
        out = pk.to_records(feat)
        assert np.array_equal(out['i'], i) and np.array_equal(out['j'], j) and np.array_equal(out['mask'] & 0x7fff, mask)
        sb, se = (feat[i] >> 15) & 1, (feat[j] >> 15) & 1
        wb, we = (feat[i] >> 14) & 1, (feat[j] >> 14) & 1
        cls = np.full(n, 7)
        cls[(sb == 0) & (se == 0)] = 0
        cls[(sb == 1) & (se == 1)] = 1
        cls[sb != se] = 2
        cls[((sb == 1) & (we == 1)) | ((se == 1) & (wb == 1))] = 3
        cls[((sb == 0) & (we == 1)) | ((se == 0) & (wb == 1))] = 4
        cls[(wb == 1) & (we == 1)] = 5
        assert np.array_equal((out['mask'] >> 16) & 7, cls)
        assert np.array_equal(out['dist'].view(np.uint32), dist.view(np.uint32))
    # a batch: the words hold j local to the structure of their row, struct_off brings the global indices back
    so = np.array([0, 700, 700, 1900, 3000], np.int32)
    n_atoms, bits = 3000, 11                                   # largest structure: 1200 atoms
    per_row = rng.integers(0, 4, size=n_atoms)
    row_off = np.concatenate([[0], np.cumsum(per_row)]).astype(np.uint32)
    i = np.repeat(np.arange(n_atoms), per_row)
    s_of = np.searchsorted(so, i, side='right') - 1
    jl = (rng.random(i.shape[0]) * (so[s_of + 1] - so[s_of])).astype(np.int64)
    mask = rng.integers(0, 1 << 15, size=i.shape[0]).astype(np.int64)
    feat = (rng.integers(0, 4, size=n_atoms).astype(np.uint32) << 14)
    pk = PackedPairs(row_off, (jl | (mask << bits)).astype(np.uint32), None, None).view(n_atoms, i.shape[0], bits, 0, False)
    out = pk.to_records(feat, struct_off=so)
    assert np.array_equal(out['i'], i) and np.array_equal(out['j'], jl + so[s_of]) and np.array_equal(out['mask'] & 0x7fff, mask)
    part = pk.structure(1900, 3000).to_records(feat[1900:3000])
    sel = i >= 1900
    assert np.array_equal(part['i'], i[sel] - 1900) and np.array_equal(part['j'], jl[sel]) and np.array_equal(part['mask'], out['mask'][sel])
