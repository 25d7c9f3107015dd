"""The device rule header compiled for the host (tests/emu) against the oracle: same pairs in, same
records out.  Checks the logic of arp_rules.cuh -- the filters, the certainly-true/false screens in
front of the exact float64 chains, the shared hydrogen loops -- without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from arpeggio_b200 import abi, params as arp_params, synth
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu():
    src = os.path.join(HERE, 'emu', 'rules_emu.cpp')
    so = os.path.join(HERE, 'emu', 'librules_emu.so')
    hdr = os.path.join(HERE, '..', 'arpeggio_b200', 'csrc', 'arp_rules.cuh')
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off', '-fno-fast-math',
                               '-o', so, src, '-lm'])
    L = C.CDLL(so)
    L.emu_classify.argtypes = [C.POINTER(abi.ArpAtoms), C.POINTER(abi.ArpParams), C.c_void_p, C.c_void_p, C.c_int64,
                               C.c_void_p, C.c_void_p, C.c_int]
    return L


def _emu_records(L, soa, p, b, e, table):
    a = soa.as_ctypes()
    b = np.ascontiguousarray(b, np.int32)
    e = np.ascontiguousarray(e, np.int32)
    out = np.zeros(b.shape[0], dtype=abi.PAIR_DTYPE)
    keep = np.zeros(b.shape[0], dtype=np.uint8)
    assert L.emu_classify(C.byref(a), C.byref(p), b.ctypes.data, e.ctypes.data, b.shape[0], out.ctypes.data,
                          keep.ctypes.data, table) == 0
    return out, keep.astype(bool)


def _within_pairs(soa, cutoff):
    from scipy.spatial import cKDTree
    pr = cKDTree(soa.xyz.astype(np.float64)).query_pairs(cutoff + 1e-3, output_type='ndarray')
    return pr[:, 0], pr[:, 1]


@pytest.mark.parametrize('case', util.golden_cases())
@pytest.mark.parametrize('table', [0, 1])
def test_emu_matches_oracle_on_golden(emu, case, table):
    g = util.Golden(case)
    n = g.soa.n_atoms
    b, e = np.triu_indices(n, 1)
    exp, keep_o = oracle.classify(g.soa, g.params, b, e)
    got, keep_e = _emu_records(emu, g.soa, g.params, b, e, table)
    assert np.array_equal(keep_o, keep_e)
    util.assert_records_equal(got[keep_e], exp[keep_o], f'{case} emu vs oracle')


@pytest.mark.parametrize('seed,cutoff,adj', [(2, 5.0, False), (5, 5.0, True), (9, 7.5, False)])
def test_emu_matches_oracle_on_clouds(emu, seed, cutoff, adj):
    soa = synth.cloud_featured(30_000, seed=seed)
    p = arp_params.make_params(cutoff, 0.1, adj)
    b, e = _within_pairs(soa, cutoff)
    exp, keep_o = oracle.classify(soa, p, b, e)
    got, keep_e = _emu_records(emu, soa, p, b, e, 1)
    assert np.array_equal(keep_o, keep_e)
    util.assert_records_equal(got[keep_e], exp[keep_o], 'cloud emu vs oracle')
    assert keep_e.sum() > 200_000


def test_emu_degenerate_hydrogens(emu):
    """Hydrogens on top of the donor / acceptor and collinear triples: the NaN -> pi fallbacks and the
    |cos| ~ 1 corner must go through the exact chain."""
    rng = np.random.default_rng(3)
    soa = synth.cloud_featured(4_000, seed=11)
    h = soa.h_xyz
    owner = np.repeat(np.arange(soa.n_atoms), np.diff(soa.h_off))
    h[::7] = soa.xyz[owner[::7]]                               # H on the donor: zero-length v1
    b, e = _within_pairs(soa, 5.0)
    # H on the acceptor of some pair / exactly collinear donor-H-acceptor
    first = {}
    for i, j in zip(b[:4000], e[:4000]):
        first.setdefault(i, j)
    for i, j in list(first.items())[:300]:
        for k in range(soa.h_off[i], soa.h_off[i + 1]):
            t = rng.choice([0.0, 0.5, 1.0, 2.0])
            h[k] = soa.xyz[i].astype(np.float64) * (1 - t) + soa.xyz[j].astype(np.float64) * t
    p = arp_params.make_params()
    exp, keep_o = oracle.classify(soa, p, b, e)
    got, keep_e = _emu_records(emu, soa, p, b, e, 1)
    assert np.array_equal(keep_o, keep_e)
    util.assert_records_equal(got[keep_e], exp[keep_o], 'degenerate hydrogens')
