"""The oracle's arithmetic models against live NumPy (the library the reference computes with):
BLAS dot/norm of 3-vectors, float32 distance, get_angle's three dtype flows, the folded plane
angle.  These pin the numerics the CUDA rules restate (arp_rules.cuh)."""
import ctypes as C
import warnings

import numpy as np
import pytest

from arpeggio_b200 import params
from oracle import oracle


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope='module')
def L():
    return oracle.lib()


def test_float32_distance_model(L):
    rng = np.random.default_rng(0)
    a = np.round(rng.uniform(-100, 100, (20000, 3)), 3).astype(np.float32)
    b = (a + rng.normal(scale=3.0, size=a.shape)).astype(np.float32)
    for x, y in zip(a, b):
        ref = np.linalg.norm(x - y)
        assert ref.dtype == np.float32
        assert L.orc_dist_f32_pub(_fp(x), _fp(y)) == ref


def test_float64_dot_and_norm_model(L):
    fma = params.probe_blas_fma()
    rng = np.random.default_rng(1)
    for _ in range(20000):
        x, y = rng.normal(size=3), rng.normal(size=3)
        assert L.orc_dot3_f64(_dp(x), _dp(y), fma) == float(np.dot(x, y))
        assert L.orc_norm3_f64(_dp(x), fma) == float(np.linalg.norm(x))


def test_float32_dot_model(L):
    rng = np.random.default_rng(2)
    for _ in range(20000):
        x, y = rng.normal(size=3).astype(np.float32), rng.normal(size=3).astype(np.float32)
        assert L.orc_dot3_f32(_fp(x), _fp(y)) == np.dot(x, y)
        assert L.orc_norm3_f32(_fp(x)) == np.linalg.norm(x)


def _get_angle(a, b, c):
    """utils.get_angle (utils.py:696-745), verbatim arithmetic."""
    v1, v2 = a - b, c - b
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        v1mag = np.sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2])
        v1norm = [v1[0] / v1mag, v1[1] / v1mag, v1[2] / v1mag]
        v2mag = np.sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2])
        v2norm = [v2[0] / v2mag, v2[1] / v2mag, v2[2] / v2mag]
        res = v1norm[0] * v2norm[0] + v1norm[1] * v2norm[1] + v1norm[2] * v2norm[2]
        angle = np.arccos(res)
    if np.isnan(angle):
        angle = np.pi
    return angle


def test_get_angle_flows(L):
    rng = np.random.default_rng(3)
    bad = 0
    n = 5000
    for _ in range(n):
        a = rng.normal(scale=2, size=3).astype(np.float32)
        c = rng.normal(scale=2, size=3).astype(np.float32)
        bd = rng.normal(scale=2, size=3)
        bf = bd.astype(np.float32)
        # arccos of the C library and of NumPy may differ in the last ulp: compare loosely here, the
        # cosine (what the CUDA path thresholds) is compared exactly through the golden fixtures
        assert abs(L.orc_get_angle_fdf(_fp(a), _dp(bd), _fp(c)) - float(_get_angle(a, bd, c))) < 1e-14
        assert abs(L.orc_get_angle_ffd(_fp(a), _fp(bf), _dp(bd + 1.0)) - float(_get_angle(a, bf, bd + 1.0))) < 1e-6
        r = _get_angle(a, bf, c)
        assert r.dtype == np.float32
        bad += abs(L.orc_get_angle_fff(_fp(a), _fp(bf), _fp(c)) - float(r)) > 3e-7
    assert bad == 0


def test_folded_plane_angle(L):
    for c in np.linspace(-1, 1, 4001):
        rad = np.arccos(np.float64(c))
        rad = rad - np.pi if rad > np.pi / 2 else rad
        assert abs(L.orc_fold_deg_f64(float(c)) - abs(rad * 180 / np.pi)) < 1e-11
        c32 = np.float32(c)
        rad = np.arccos(c32)
        rad = rad - np.pi if rad > np.pi / 2 else rad
        ref = abs(rad * 180 / np.pi)
        assert ref.dtype == np.float32
        assert abs(L.orc_fold_deg_f32(float(c32)) - float(ref)) < 2e-5


def test_cosine_images_reproduce_the_angle_tests():
    """Thresholding the cosine with the bisected images == thresholding arccos, on dense samples
    around every edge (params.py)."""
    p = params.make_params()
    f8 = np.float64
    for thr, img in ((p.hbond_angle, p.cos_hbond), (p.weak_hbond_angle, p.cos_weak_hbond), (p.cx_angle_min, p.cos_cx_min)):
        c = f8(img)
        for _ in range(200):
            assert (np.arccos(c) >= thr) == (c <= img)
            c = np.nextafter(c, f8(2))
        c = f8(img)
        for _ in range(200):
            assert (np.arccos(c) >= thr) == (c <= img)
            c = np.nextafter(c, f8(-2))
    c = np.float32(p.cos_xbond_f32)
    for step in (np.float32(2), np.float32(-2)):
        x = c
        for _ in range(200):
            assert (np.arccos(x) >= p.xbond_angle) == (x <= np.float32(p.cos_xbond_f32))
            x = np.nextafter(x, step)
    for ft, suffix in ((np.float64, 'f64'), (np.float32, 'f32')):
        split = ft(getattr(p, 'cos_split_' + suffix))
        pos, neg = getattr(p, 'cos_pos_' + suffix), getattr(p, 'cos_neg_' + suffix)
        for k in range(3):
            b = p.plane_bins_deg[k]
            for centre in (ft(pos[k]), ft(neg[k]), split):
                if abs(centre) > 1:
                    continue
                for step in (ft(2), ft(-2)):
                    x = centre
                    for _ in range(100):
                        if abs(x) <= 1:
                            want = params._fold_deg(x) <= b
                            got = (x <= ft(neg[k])) if x <= split else (x >= ft(pos[k]))
                            assert want == got, (suffix, k, float(x))
                        x = np.nextafter(x, step)
