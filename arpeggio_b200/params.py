"""Run-time parameters of the contact kernels.

The numeric thresholds are those of ``arpeggio.core.config`` (config.py:23-25,
:592-660); a host that has the reference importable passes its live
``config.CONTACT_TYPES`` so that ``config`` stays the single source of truth.

The GPU never evaluates ``arccos``.  Every angle test of the reference
(utils.py:90, :113, :151, :176; interactions.py:1007, :1129-1148, :1282, :1363)
is a comparison of ``arccos(c)`` (possibly folded and converted to degrees) with
a constant, and ``arccos`` is monotone, so each test is equivalent to comparing
the cosine ``c`` with the image of that constant.  The images are found here by
bisection over the bit patterns of float64 / float32 using the HOST's NumPy
``arccos`` and the reference's exact expression, which makes the kernel agree
with whatever libm/SVML the host NumPy dispatches to, bit for bit.
"""
import warnings

import numpy as np

from . import abi

# config.CONTACT_TYPES (config.py:592-660) and config.VDW_RADII (config.py:23-25)
DEFAULT_CONTACT_TYPES = {
    'hbond': {'distance': 3.9, 'polar distance': 3.5, 'angle rad': 1.57},
    'weak hbond': {'distance': 3.6, 'weak polar distance': 3.5, 'angle rad': 2.27,
                   'cx angle min rad': 0.52, 'cx angle max rad': 2.62},
    'aromatic': {'distance': 4.0, 'centroid_distance': 6.0, 'atom_aromatic_distance': 4.5,
                 'met_sulphur_aromatic_distance': 6.0},
    'amide': {'centroid_distance': 6.0, 'angle degree': 30.0},
    'xbond': {'angle theta 1 rad': 2.09},
    'ionic': {'distance': 4.0},
    'hydrophobic': {'distance': 4.5},
    'carbonyl': {'distance': 3.6},
    'metal': {'distance': 2.8},
}
DEFAULT_DIST_MAX = 4.5      # config.CONTACT_TYPES_DIST_MAX
DEFAULT_H_VDW = 1.2         # config.VDW_RADII['H']
PLANE_BINS_DEG = (30.0, 60.0, 90.0)   # literals at interactions.py:1129-1148, :1007, :1282, :1363


# ---------------------------------------------------------------------------
# ordered-integer views of floats for bisection
# ---------------------------------------------------------------------------
def _ord(x, ft, it):
    b = int(np.array(x, dtype=ft).view(it))
    sign = 1 << (8 * np.dtype(it).itemsize - 1)
    return -(b & (sign - 1)) if b < 0 else b


def _unord(k, ft, it):
    sign = 1 << (8 * np.dtype(it).itemsize - 1)
    b = k if k >= 0 else (-k) | sign
    if b >= sign:
        b -= 2 * sign
    return np.array(b, dtype=it).view(ft)[()]


def _edge(pred, lo, hi, ft, it, true_below):
    """Bisect the single switch of a monotone predicate over the floats [lo, hi].

    true_below: pred is True for small arguments.  Returns the last float for which
    pred is True on its True side (largest if true_below else smallest), or None if
    pred is constant over the interval.
    """
    klo, khi = _ord(lo, ft, it), _ord(hi, ft, it)
    plo, phi = bool(pred(_unord(klo, ft, it))), bool(pred(_unord(khi, ft, it)))
    if plo == phi:
        return None
    if plo != true_below:
        raise ValueError('predicate orientation differs from the expected one')
    while khi - klo > 1:
        mid = (klo + khi) // 2
        if bool(pred(_unord(mid, ft, it))) == plo:
            klo = mid
        else:
            khi = mid
    edge = klo if true_below else khi
    # the switch must be clean in a window around the edge (guards against a non-monotone libm)
    for k in range(max(edge - 48, _ord(lo, ft, it)), min(edge + 48, _ord(hi, ft, it)) + 1):
        expect = (k <= edge) if true_below else (k >= edge)
        if bool(pred(_unord(k, ft, it))) != expect:
            warnings.warn('arccos is not monotone near an angle threshold; the cosine image is approximate')
            break
    return _unord(edge, ft, it)


def _fold_deg(c):
    """|group_angle(..., degrees=True, signed=True)| for a NumPy scalar cosine (utils.py:646-660)."""
    rad = np.arccos(c)
    rad = rad - np.pi if rad > np.pi / 2 else rad
    return abs(rad * 180 / np.pi)


def cosine_images(p):
    """Fill the cos_* members of an ArpParams from its angle members, with the host NumPy."""
    f8, i8, f4, i4 = np.float64, np.int64, np.float32, np.int32
    one, mone = f8(1.0), f8(-1.0)

    def ge(thr):   # largest c with arccos(c) >= thr  (utils.py:90 `get_angle(...) >= angle`)
        e = _edge(lambda c: np.arccos(c) >= thr, mone, one, f8, i8, True)
        if e is None:
            return 2.0 if np.arccos(one) >= thr else -2.0
        return float(e)

    p.cos_hbond = ge(p.hbond_angle)
    p.cos_weak_hbond = ge(p.weak_hbond_angle)
    p.cos_cx_min = ge(p.cx_angle_min)          # utils.py:151  min <= angle
    e = _edge(lambda c: np.arccos(c) <= p.cx_angle_max, mone, one, f8, i8, False)   # angle <= max
    p.cos_cx_max = float(e) if e is not None else (-2.0 if np.arccos(mone) <= p.cx_angle_max else 2.0)
    # is_xbond: float32 theta >= python float (utils.py:176) -> float32 comparison under NEP 50
    thr = p.xbond_angle
    e = _edge(lambda c: np.arccos(c) >= thr, f4(-1.0), f4(1.0), f4, i4, True)
    p.cos_xbond_f32 = float(e) if e is not None else (2.0 if np.arccos(f4(1.0)) >= thr else -2.0)

    for ft, it, suffix in ((f8, i8, 'f64'), (f4, i4, 'f32')):
        lo, hi = ft(-1.0), ft(1.0)
        split = _edge(lambda c: np.arccos(c) > np.pi / 2, lo, hi, ft, it, True)
        assert split is not None
        nxt = _unord(_ord(split, ft, it) + 1, ft, it)
        pos = getattr(p, 'cos_pos_' + suffix)
        neg = getattr(p, 'cos_neg_' + suffix)
        for k in range(3):
            b = p.plane_bins_deg[k]
            e = _edge(lambda c: _fold_deg(c) <= b, nxt, hi, ft, it, False)          # pos branch: c >= edge
            pos[k] = float(e) if e is not None else (-2.0 if _fold_deg(hi) <= b else 2.0)
            e = _edge(lambda c: _fold_deg(c) <= b, lo, split, ft, it, True)         # neg branch: c <= edge
            neg[k] = float(e) if e is not None else (2.0 if _fold_deg(lo) <= b else -2.0)
        setattr(p, 'cos_split_' + suffix, float(split))
    return p


def probe_blas_fma():
    """1 if np.dot on float64 3-vectors is the FMA chain of OpenBLAS' Haswell/SkylakeX ddot, 0 if
    it is the plain sequential sum.  Decided on vectors where the two differ."""
    rng = np.random.default_rng(12345)
    votes = [0, 0]
    for _ in range(400):
        x = rng.normal(size=3)
        y = rng.normal(size=3)
        seq = float((x[0] * y[0] + x[1] * y[1]) + x[2] * y[2])
        # exact FMA through integer arithmetic on the 53-bit significands
        fm = _fma(x[2], y[2], _fma(x[1], y[1], float(x[0] * y[0])))
        if seq == fm:
            continue
        d = float(np.dot(x, y))
        if d == fm:
            votes[1] += 1
        elif d == seq:
            votes[0] += 1
    if votes[0] and votes[1]:
        warnings.warn('np.dot follows neither the sequential nor the FMA model consistently')
    return 1 if votes[1] >= votes[0] else 0


def _fma(a, b, c):
    from fractions import Fraction
    r = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    return float(r)   # Fraction -> float is correctly rounded


_BLAS_FMA = None


def make_params(interacting_cutoff=5.0, vdw_comp=0.1, include_sequence_adjacent=False,
                contact_types=None, dist_max=DEFAULT_DIST_MAX, h_vdw=DEFAULT_H_VDW, blas_fma=None):
    """ArpParams for one run_arpeggio(..) call (interactions.py:329-347)."""
    global _BLAS_FMA
    ct = contact_types or DEFAULT_CONTACT_TYPES
    p = abi.ArpParams()
    p.interacting_cutoff = float(interacting_cutoff)
    p.vdw_comp = float(vdw_comp)
    p.include_sequence_adjacent = 1 if include_sequence_adjacent else 0
    if blas_fma is None:
        if _BLAS_FMA is None:
            _BLAS_FMA = probe_blas_fma()
        blas_fma = _BLAS_FMA
    p.blas_fma = int(blas_fma)
    p.h_vdw = float(h_vdw)
    p.dist_max = float(dist_max)
    p.hbond_polar_dist = ct['hbond']['polar distance']
    p.weak_polar_dist = ct['weak hbond']['weak polar distance']
    p.ionic_dist = ct['ionic']['distance']
    p.carbonyl_dist = ct['carbonyl']['distance']
    p.aromatic_dist = ct['aromatic']['distance']
    p.hydrophobic_dist = ct['hydrophobic']['distance']
    p.metal_dist = ct['metal']['distance']
    p.hbond_angle = ct['hbond']['angle rad']
    p.weak_hbond_angle = ct['weak hbond']['angle rad']
    p.cx_angle_min = ct['weak hbond']['cx angle min rad']
    p.cx_angle_max = ct['weak hbond']['cx angle max rad']
    p.xbond_angle = ct['xbond']['angle theta 1 rad']
    p.ring_centroid_dist = ct['aromatic']['centroid_distance']
    p.atom_ring_dist = ct['aromatic']['atom_aromatic_distance']
    p.met_sulphur_dist = ct['aromatic']['met_sulphur_aromatic_distance']
    p.amide_centroid_dist = ct['amide']['centroid_distance']
    for k in range(3):
        p.plane_bins_deg[k] = PLANE_BINS_DEG[k]
    key = _cache_key(p)
    img = _IMAGE_CACHE.get(key)
    if img is None:
        cosine_images(p)
        _IMAGE_CACHE[key] = _snapshot(p)
    else:
        _restore(p, img)
    return p


_IMAGE_CACHE = {}
_IMG_FIELDS = ('cos_hbond', 'cos_weak_hbond', 'cos_cx_min', 'cos_cx_max', 'cos_xbond_f32', 'cos_split_f64',
               'cos_split_f32')
_IMG_ARRAYS = ('cos_pos_f64', 'cos_neg_f64', 'cos_pos_f32', 'cos_neg_f32')


def _cache_key(p):
    return (p.hbond_angle, p.weak_hbond_angle, p.cx_angle_min, p.cx_angle_max, p.xbond_angle,
            tuple(p.plane_bins_deg))


def _snapshot(p):
    return ({f: getattr(p, f) for f in _IMG_FIELDS}, {f: list(getattr(p, f)) for f in _IMG_ARRAYS})


def _restore(p, img):
    for f, v in img[0].items():
        setattr(p, f, v)
    for f, v in img[1].items():
        a = getattr(p, f)
        for k in range(3):
            a[k] = v[k]
