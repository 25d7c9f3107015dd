"""GPU parity of the atom-atom path: libarpeggio_cuda.so (through the C ABI) against the CPU oracle
and against the golden records produced by the reference's own code.  Bit-exact: pair set, (i < j)
orientation, 15-bit mask + entity class, float32 distance bits."""
import dataclasses

import numpy as np
import pytest

import util
from arpeggio_b200 import abi, params as arp_params, synth
from arpeggio_b200.soa import AtomSoA
from oracle import oracle

pytestmark = pytest.mark.gpu

CASES = util.golden_cases()


@pytest.mark.parametrize('case', CASES)
def test_golden_pairs(engine, case):
    g = util.Golden(case)
    engine.set_params(g.params)
    got = engine.pairs(g.soa)
    if g.meta['raises'] == 'AttributeError':
        assert np.any(got['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR))
        util.assert_records_equal(got, oracle.pairs(g.soa, g.params), f'{case} vs oracle')
        return
    util.assert_records_equal(got, g.exp_pairs, f'{case} atom-atom vs reference records')


def _check_vs_oracle(engine, soa, p, what):
    engine.set_params(p)
    got = engine.pairs(soa)
    exp, n_within = oracle.pairs(soa, p, with_counts=True)
    util.assert_records_equal(got, exp, what)
    st = engine.stats()
    assert st['n_pairs'] == exp.shape[0]
    assert st['n_candidates'] >= n_within
    # the unsorted stream is a permutation of the sorted one
    raw = engine.fetch_pairs(got.shape[0], sorted=False)
    util.assert_records_equal(util.sort_pairs(raw), got, what + ' (unsorted stream)')
    return got


def test_config2_uniform_cloud(engine):
    soa = synth.cloud_uniform(10_000, seed=1)
    got = _check_vs_oracle(engine, soa, arp_params.make_params(), 'config 2')
    assert got.shape[0] > 100_000
    assert not np.any(got['mask'] & np.uint32(0x7FE0))        # bits 5..14 need feature masks


def test_config3_featured_cloud_full_size(engine):
    soa = synth.cloud_featured(100_000, seed=2)
    got = _check_vs_oracle(engine, soa, arp_params.make_params(), 'config 3')
    assert got.shape[0] > 1_000_000
    seen = np.bitwise_or.reduce(got['mask'])
    assert seen & 0x7FFF == 0x7FFF, 'every SIFt bit occurs in the 100k cloud'
    assert not np.any(got['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR))


@pytest.mark.parametrize('cutoff,comp,adj', [(3.0, 0.1, False), (5.0, 0.0, True), (8.0, 0.35, False), (0.5, 0.1, False)])
def test_parameter_sweep(engine, cutoff, comp, adj):
    soa = synth.cloud_featured(6_000, seed=7)
    _check_vs_oracle(engine, soa, arp_params.make_params(cutoff, comp, adj), f'cutoff={cutoff}')


@pytest.mark.parametrize('seed', range(12))
def test_randomized_structures(engine, seed):
    """Seeded sweep over size, density, cutoff, compensation factor, hydrogen bond lengths (the hydrogen-reach screen
    of k_classify depends on the longest one) and residue sizes; every stream bit-identical to the oracle's, and so
    are the per-atom SIFt reductions."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([37, 300, 2_000, 9_000, 25_000]))
    soa = synth.cloud_featured(n, seed=2000 + seed, atoms_per_residue=int(rng.integers(1, 12)), chain_len=int(rng.integers(2, 50)))
    scale = float(rng.choice([0.6, 1.0, 1.0, 1.7]))             # denser / sparser than protein density
    soa.xyz[:] = np.round(soa.xyz.astype(np.float64) * scale, 3).astype(np.float32)
    owner = np.repeat(np.arange(n), np.diff(soa.h_off))
    d = rng.normal(size=(owner.shape[0], 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    length = rng.uniform(0.5, float(rng.choice([1.1, 1.1, 2.5])), size=(owner.shape[0], 1))
    soa.h_xyz[:] = soa.xyz[owner].astype(np.float64) + d * length
    if soa.xnbr_xyz is not None:
        soa.xnbr_xyz[:] = (soa.xnbr_xyz.astype(np.float64) * scale).astype(np.float32)
    p = arp_params.make_params(float(rng.choice([4.0, 5.0, 5.0, 6.5])), float(rng.choice([0.0, 0.1, 0.1, 0.4])), bool(rng.integers(2)))
    got = _check_vs_oracle(engine, soa, p, f'seed {seed}: n={n} scale={scale}')
    sifts, exp = engine.atom_sifts(), oracle.atom_sifts(got, n)
    for f in sifts.dtype.names:
        assert np.array_equal(sifts[f], exp[f]), f


def test_far_from_origin(engine):
    soa = synth.cloud_featured(5_000, seed=8)
    soa.xyz += np.float32(9000.0)          # large coordinates widen the float32 prefilter band
    soa.h_xyz += 9000.0
    soa.xnbr_xyz += np.float32(9000.0)
    _check_vs_oracle(engine, soa, arp_params.make_params(), 'offset cloud')


def test_degenerate_single_cell(engine):
    """Hundreds of atoms inside one cell (candidate chunks > 128, long home lists)."""
    rng = np.random.default_rng(5)
    n = 700
    soa = synth.cloud_featured(n, seed=9)
    soa.xyz[:] = np.round(rng.uniform(0, 4.0, size=(n, 3)), 3).astype(np.float32)
    soa.xyz[:40] = soa.xyz[0]               # coincident atoms (distance 0)
    _check_vs_oracle(engine, soa, arp_params.make_params(), 'degenerate')


def test_line_and_plane_shapes(engine):
    """Elongated bounding boxes: the cell bound (4 n + 64) forces wider cells."""
    rng = np.random.default_rng(6)
    n = 3000
    soa = synth.cloud_featured(n, seed=10)
    soa.xyz[:, 0] = np.round(rng.uniform(0, 30000.0, n), 3)
    soa.xyz[:, 1:] = np.round(rng.uniform(0, 3.0, (n, 2)), 3)
    _check_vs_oracle(engine, soa, arp_params.make_params(), 'line')
    soa.xyz[:, 0] = np.round(rng.uniform(0, 150.0, n), 3)
    _check_vs_oracle(engine, soa, arp_params.make_params(), 'rod')


def test_knife_edge_distances(engine):
    """Pairs whose float32 distance sits exactly on, one ulp below and one ulp above each threshold."""
    p = arp_params.make_params()
    thr = [5.0, 4.5, 4.0, 3.9, 3.6, 3.5, 3.4, 3.3, 2.8, 1.52, 1.42]
    xs, feats = [], []
    x0 = 0.0
    for t in thr:
        f = np.float32(t)
        for d in (np.nextafter(f, np.float32(0)), f, np.nextafter(f, np.float32(10))):
            xs += [x0, x0 + float(d)]
            x0 += 40.0
    n = len(xs)
    xyz = np.zeros((n, 3), np.float32)
    xyz[:, 0] = np.array(xs, dtype=np.float64).astype(np.float32)
    all_types = np.uint32(0xFFF) | abi.F_IS_METAL
    soa = AtomSoA(xyz=xyz, feat=np.full(n, all_types, np.uint32), res_id=np.arange(n, dtype=np.int32),
                  rad_class=(np.arange(n) % 2).astype(np.uint16), vdw=np.array([1.7, 1.6]), cov=np.array([0.76, 0.66]),
                  res_prev=np.full(n, -1, np.int32), res_next=np.full(n, -1, np.int32), res_flags=np.zeros(n, np.uint8))
    got = _check_vs_oracle(engine, soa, p, 'knife edge')
    assert got.shape[0] >= 2 * len(thr)


def test_empty_and_tiny(engine):
    p = arp_params.make_params()
    engine.set_params(p)
    for n in (0, 1, 2):
        soa = synth.cloud_featured(8, seed=3)
        sub = AtomSoA(xyz=soa.xyz[:n], feat=soa.feat[:n], res_id=np.arange(n, dtype=np.int32), rad_class=soa.rad_class[:n],
                      vdw=soa.vdw, cov=soa.cov, res_prev=np.full(max(n, 1), -1, np.int32),
                      res_next=np.full(max(n, 1), -1, np.int32), res_flags=np.zeros(max(n, 1), np.uint8))
        got = engine.pairs(sub)
        util.assert_records_equal(got, oracle.pairs(sub, p), f'n={n}')


def test_batch_equals_per_structure(engine):
    """A batch (struct_off) gives exactly the union of the per-structure streams; no pair spans two structures."""
    p = arp_params.make_params()
    engine.set_params(p)
    parts = [synth.cloud_featured(n, seed=20 + k) for k, n in enumerate((1500, 1, 0, 3000, 700))]
    parts[2] = AtomSoA(xyz=np.zeros((0, 3), np.float32), feat=np.zeros(0, np.uint32), res_id=np.zeros(0, np.int32),
                       rad_class=np.zeros(0, np.uint16), vdw=parts[0].vdw, cov=parts[0].cov,
                       res_prev=np.zeros(0, np.int32), res_next=np.zeros(0, np.int32), res_flags=np.zeros(0, np.uint8))
    batch = AtomSoA.concat(parts)
    got = engine.pairs(batch)
    util.assert_records_equal(got, oracle.pairs(batch, p), 'batch vs oracle')
    off = batch.struct_off
    s_i = np.searchsorted(off, got['i'], side='right')
    s_j = np.searchsorted(off, got['j'], side='right')
    assert np.array_equal(s_i, s_j)
    pieces = []
    for k, part in enumerate(parts):
        if part.n_atoms == 0:
            continue
        r = engine.pairs(part).copy()
        r['i'] += off[k]
        r['j'] += off[k]
        pieces.append(r)
    util.assert_records_equal(got, util.sort_pairs(np.concatenate(pieces)), 'batch vs per-structure runs')


@pytest.mark.parametrize('knob', [None, 'ARPEGGIO_NO_REG_GRID', 'ARPEGGIO_NO_FUSED_GRID', 'ARPEGGIO_NO_PDL',
                                  'ARPEGGIO_NO_EARLY_CLASSIFY', 'ARPEGGIO_TILES'])
def test_every_grid_build_path(monkeypatch, knob):
    """The cell grid is built by one of three code paths (register-cached cooperative kernel with 1 or 2 atoms
    per thread, cooperative kernel through global memory, five kernels); each must give the oracle's stream for
    a single structure of either size class and for a small batch.  ARPEGGIO_TILES: the fused search + classify tile
    kernel (k_tiles) in place of k_search + k_classify, plus a dense structure (cells with several jobs) and a golden
    fixture of the reference."""
    from arpeggio_b200.engine import ContactEngine
    if knob:
        monkeypatch.setenv(knob, '1')
    p = arp_params.make_params()
    parts = [synth.cloud_featured(n, seed=70 + k) for k, n in enumerate((4000, 9, 2500))]
    cases = {'30k': synth.cloud_featured(30_000, seed=61), '200k': synth.cloud_featured(200_000, seed=62),
             'batch': AtomSoA.concat(parts)}
    if knob == 'ARPEGGIO_NO_PDL':
        del cases['200k']
    if knob == 'ARPEGGIO_TILES':
        dense = synth.cloud_featured(4_000, seed=63)
        dense.xyz[:] = np.round(dense.xyz.astype(np.float64) * 0.45, 3).astype(np.float32)     # ~11x protein density
        dense.h_xyz[:] = dense.h_xyz * 0.45
        cases['dense'] = dense
    with ContactEngine(0, p) as eng:
        for name, soa in cases.items():
            util.assert_records_equal(eng.pairs(soa), oracle.pairs(soa, p), f'{knob or "default"} {name}')
        if knob == 'ARPEGGIO_TILES':
            g = util.Golden('rich_whole')
            eng.set_params(g.params)
            util.assert_records_equal(eng.pairs(g.soa), g.exp_pairs, 'tiles: golden rich_whole')


def test_skewed_density_and_odd_sizes():
    """Inputs that stress the hand-over between k_search and k_classify and the capacity logic: three atoms, a batch
    with an empty structure, and one structure whose dense corner holds almost all candidates while a sparse remainder
    stretches the grid (widened cells; the first guess of the list capacity overflows and is regrown)."""
    from arpeggio_b200.engine import ContactEngine
    p = arp_params.make_params()
    parts = [synth.cloud_featured(n, seed=170 + k) for k, n in enumerate((3000, 0, 17, 2500))]
    dense = synth.cloud_featured(20_000, seed=163)
    far = synth.cloud_featured(20_000, seed=164)
    xyz = np.concatenate([dense.xyz * np.float32(0.6), far.xyz * np.float32(3.0) + np.float32(400.0)])
    skew = dataclasses.replace(AtomSoA.concat([dense, far]), xyz=xyz, struct_off=None)
    cases = {'60k': synth.cloud_featured(60_000, seed=161), '3 atoms': synth.cloud_featured(3, seed=162),
             'batch': AtomSoA.concat(parts), 'skewed': skew}
    with ContactEngine(0, p) as eng:
        for name, soa in cases.items():
            util.assert_records_equal(eng.pairs(soa), oracle.pairs(soa, p), name)
        util.assert_records_equal(eng.pairs(cases['60k']), oracle.pairs(cases['60k'], p), '60k after the regrown run')


@pytest.mark.parametrize('case', [c for c in CASES if c != 'xbond_fault'])
def test_atom_sifts_golden(engine, case):
    """atom.sift*, integer_sift*, actual_hbonds*, actual_polars* as the reference left them."""
    g = util.Golden(case)
    engine.set_params(g.params)
    engine.pairs(g.soa)
    got = engine.atom_sifts()
    util.assert_atom_sifts_equal(got, g, case)


def test_atom_sifts_full_size(engine):
    p = arp_params.make_params()
    engine.set_params(p)
    soa = synth.cloud_featured(100_000, seed=2)
    rec = engine.pairs(soa)
    got = engine.atom_sifts()
    exp = oracle.atom_sifts(rec, soa.n_atoms)
    for f in got.dtype.names:
        assert np.array_equal(got[f], exp[f]), f
    # size-independent properties: every record touches two atoms; OR of the atom words = OR of the record masks
    assert int(got['hbonds'][:, 0].sum()) == 2 * int(np.count_nonzero(rec['mask'] >> 5 & 1))
    assert int(got['polars'][:, 0].sum()) == 2 * int(np.count_nonzero(rec['mask'] >> 13 & 1))
    assert np.bitwise_or.reduce(got['sift'][:, 0]) == np.bitwise_or.reduce(rec['mask']) & 0x7FFF
    assert np.array_equal(got['sift'][:, 0], got['sift'][:, 1] | got['sift'][:, 2] | got['sift'][:, 3])
    # an empty structure and a structure without contacts
    empty = AtomSoA(xyz=np.zeros((0, 3), np.float32), feat=np.zeros(0, np.uint32), res_id=np.zeros(0, np.int32),
                    rad_class=np.zeros(0, np.uint16), vdw=soa.vdw, cov=soa.cov, res_prev=np.zeros(0, np.int32),
                    res_next=np.zeros(0, np.int32), res_flags=np.zeros(0, np.uint8))
    engine.pairs(empty)
    assert engine.atom_sifts().shape == (0,)
    lone = synth.cloud_featured(3, seed=5)
    lone.xyz[:] = np.array([[0, 0, 0], [50, 0, 0], [0, 50, 0]], np.float32)
    assert engine.pairs(lone).shape[0] == 0
    z = engine.atom_sifts()
    assert z.shape == (3,) and not z['sift'].any() and not z['integer_sift'].any()


def test_single_block_upload_equals_per_array_upload(engine):
    """engine.pinned_soa lays a structure out as one pinned block that goes up in one DMA (device arrays = views of
    an arena); results must not depend on the upload path, also when the two alternate or the arena has to grow."""
    from arpeggio_b200.engine import pinned_soa
    p = arp_params.make_params()
    engine.set_params(p)
    small, big = synth.cloud_featured(3000, seed=81), synth.cloud_featured(40_000, seed=82)
    parts = [synth.cloud_featured(n, seed=90 + k) for k, n in enumerate((2000, 5, 1200))]
    batch = AtomSoA.concat(parts)
    exp = {id(s): oracle.pairs(s, p) for s in (small, big, batch)}
    for soa in (small, pinned_soa(small), pinned_soa(big), small, pinned_soa(batch), big, pinned_soa(small)):
        key = [s for s in (small, big, batch) if s.n_atoms == soa.n_atoms][0]
        util.assert_records_equal(engine.pairs(soa), exp[id(key)], f'upload of {soa.n_atoms} atoms')
    for f in ('sift', 'integer_sift'):
        assert np.array_equal(engine.atom_sifts()[f], oracle.atom_sifts(exp[id(small)], small.n_atoms)[f])
    nohyd = AtomSoA(xyz=small.xyz, feat=small.feat, res_id=small.res_id, rad_class=small.rad_class, vdw=small.vdw,
                    cov=small.cov, res_prev=small.res_prev, res_next=small.res_next, res_flags=small.res_flags)
    util.assert_records_equal(engine.pairs(pinned_soa(nohyd)), oracle.pairs(nohyd, p), 'no hydrogens: every scan screened out')


def test_rerun_is_stable_and_overflow_regrows(engine):
    """A sparse structure sizes the record buffer small; a dense one must regrow it (overflow path)."""
    p = arp_params.make_params()
    engine.set_params(p)
    from arpeggio_b200.engine import ContactEngine
    with ContactEngine(engine.device, p) as e2:
        sparse = synth.cloud_uniform(200, seed=4)
        sparse.xyz *= np.float32(10.0)
        assert e2.pairs(sparse).shape[0] < 50
        rng = np.random.default_rng(1)
        dense = synth.cloud_uniform(2000, seed=5)
        dense.xyz[:] = np.round(rng.uniform(0, 12.0, size=(2000, 3)), 3).astype(np.float32)
        got = e2.pairs(dense)
        assert got.shape[0] > 2000 * 16 + 4096
        util.assert_records_equal(got, oracle.pairs(dense, p), 'dense after sparse')
        util.assert_records_equal(e2.pairs(), got, 'second run on resident inputs')


def test_early_and_late_classify_alternate():
    """Up to 5 * 10^5 atoms k_classify starts on the candidates while k_search drains and leaves the candidate list
    zeroed; larger inputs wait for k_search and leave the list as it is.  Alternating the two on one context must
    keep every stream identical to the oracle's (the list is zeroed again before an early run)."""
    from arpeggio_b200.engine import ContactEngine
    p = arp_params.make_params()
    small = synth.cloud_featured(50_000, seed=301)
    large = synth.cloud_featured(520_000, seed=302)
    exp_small = oracle.pairs(small, p)
    with ContactEngine(0, p) as eng:
        util.assert_records_equal(eng.pairs(small), exp_small, 'early start, first run')
        util.assert_records_equal(eng.pairs(large), oracle.pairs(large, p), 'late start, 520k atoms')
        for k in range(3):
            util.assert_records_equal(eng.pairs(small), exp_small, f'early start after a late one, run {k}')


def test_flag_within(engine):
    soa = synth.cloud_featured(20_000, seed=11)
    soa.feat &= ~np.uint32(abi.F_IN_SELECTION)
    soa.feat[100:160] |= abi.F_IN_SELECTION
    engine.upload_atoms(soa)
    for radius in (6.0, 0.0, 12.5):
        assert np.array_equal(engine.flag_within(radius), oracle.flag_within(soa, radius)), radius


def test_flag_within_grid_path(engine, monkeypatch):
    """Large inputs take the cell-grid variant (>= 20 000 atoms): a big selection (a 'chain' of 12 000 atoms), a batch of
    structures that overlap in space (flags never cross structures), atoms far outside the bulk, a radius larger than
    the box; the double loop (knob) gives the same flags."""
    from arpeggio_b200.engine import ContactEngine
    soa = synth.cloud_featured(60_000, seed=12)
    soa.feat &= ~np.uint32(abi.F_IN_SELECTION)
    soa.feat[5_000:17_000] |= abi.F_IN_SELECTION
    soa.xyz[:50] += np.float32(400.0)                   # outliers: clamped into border cells
    soa.xyz[59_000:59_020] -= np.float32(250.0)
    soa.feat[59_010] |= abi.F_IN_SELECTION
    engine.upload_atoms(soa)
    exp = {r: oracle.flag_within(soa, r) for r in (6.0, 0.0, 2.5, 300.0)}
    for r, e in exp.items():
        assert np.array_equal(engine.flag_within(r), e), r
    parts = [synth.cloud_featured(n, seed=40 + k) for k, n in enumerate((9_000, 14_000, 30, 8_000))]
    for k, part in enumerate(parts):
        part.feat &= ~np.uint32(abi.F_IN_SELECTION)
        part.feat[k * 7:k * 7 + 40] |= abi.F_IN_SELECTION
    batch = AtomSoA.concat(parts)
    engine.upload_atoms(batch)
    assert np.array_equal(engine.flag_within(6.0), oracle.flag_within(batch, 6.0))
    monkeypatch.setenv('ARPEGGIO_NO_WITHIN_GRID', '1')
    with ContactEngine(engine.device) as e2:
        e2.upload_atoms(soa)
        assert np.array_equal(e2.flag_within(6.0), exp[6.0])


def test_errors_are_reported(engine):
    from arpeggio_b200._lib import ArpeggioCudaError
    from arpeggio_b200.engine import ContactEngine
    with ContactEngine(engine.device) as e2:
        with pytest.raises(ArpeggioCudaError) as ei:
            e2.run_pairs()
        assert ei.value.code == abi.E_NOT_READY
        soa = synth.cloud_uniform(10, seed=1)
        bad = AtomSoA(xyz=soa.xyz.copy(), feat=soa.feat, res_id=soa.res_id, rad_class=soa.rad_class, vdw=soa.vdw, cov=soa.cov,
                      res_prev=soa.res_prev, res_next=soa.res_next, res_flags=soa.res_flags)
        bad.xyz[3, 1] = np.nan
        with pytest.raises(ValueError):
            e2.upload_atoms(bad)
        e2.upload_atoms(soa)
        n = e2.run_pairs()
        small = np.empty(max(n - 1, 0), dtype=abi.PAIR_DTYPE)
        if n:
            with pytest.raises(ArpeggioCudaError) as ei:
                e2._check(e2._L.arp_pairs_fetch(e2._ctx, small.ctypes.data, small.shape[0], 1))
            assert ei.value.code == abi.E_CAPACITY
    with pytest.raises(ArpeggioCudaError):
        ContactEngine(device=4096)


@pytest.mark.parametrize('n', [0, 1, 40, 3000, 60_000])
def test_compact_stream_and_async_run(engine, n):
    """arp_pairs_run_async + arp_pairs_fetch_compact: the compact view (row offsets, (j, mask) records, distances as a
    stream of their own) unpacks to exactly the sorted 16-byte stream; the fetch is what waits for the run."""
    p = arp_params.make_params()
    engine.set_params(p)
    soa = synth.cloud_featured(max(n, 1), seed=31)
    if n == 0:
        soa = AtomSoA(xyz=soa.xyz[:0], feat=soa.feat[:0], res_id=soa.res_id[:0], rad_class=soa.rad_class[:0], vdw=soa.vdw,
                      cov=soa.cov, res_prev=soa.res_prev, res_next=soa.res_next, res_flags=soa.res_flags)
    engine.upload_atoms(soa)
    engine.run_pairs_async()
    cp = engine.fetch_pairs_compact(with_dist=True)            # waits for the run, sizes its own buffers
    exp = engine.fetch_pairs(cp.n, sorted=True)
    assert cp.n == exp.shape[0] == engine.pair_count()
    assert cp.row_off.shape[0] == soa.n_atoms + 1 and int(cp.row_off[-1]) == cp.n
    util.assert_records_equal(cp.to_records(), exp, f'compact n={n}')
    assert np.all(np.diff(cp.row_off.astype(np.int64)) >= 0)
    # the split stream: records now, distances later
    engine.run_pairs_async()
    cp2 = engine.fetch_pairs_compact(with_dist=False)
    assert cp2.dist is None and cp2.nbytes == 4 * (soa.n_atoms + 1) + 8 * cp2.n
    d = engine.fetch_pairs_dist(cp2.n)
    util.assert_records_equal(cp2.to_records(dist=d), exp, f'compact + distances on demand n={n}')
    if n >= 3000:
        util.assert_records_equal(exp, oracle.pairs(soa, p), 'against the oracle')


def test_capacity_error_reports_the_count(engine):
    """A destination that is too small fails with ARP_E_CAPACITY, nothing is written, the count comes back."""
    import ctypes as C
    engine.set_params(arp_params.make_params())
    soa = synth.cloud_featured(2000, seed=32)
    engine.upload_atoms(soa)
    n = engine.run_pairs()
    row = np.zeros(soa.n_atoms + 1, np.uint32)
    rec = np.full(8, 7, dtype=np.int64).view(abi.PAIR_C_DTYPE)
    got = C.c_uint64()
    rc = engine._L.arp_pairs_fetch_compact(engine._ctx, row.ctypes.data, rec.ctypes.data, rec.shape[0], None, C.byref(got))
    assert rc == abi.E_CAPACITY and got.value == n and np.all(rec.view(np.int64) == 7)
    assert engine.stats()['faults'] == 0


@pytest.mark.parametrize('n', [0, 3, 5000, 140_000])
def test_packed_stream(engine, n):
    """arp_pairs_fetch_packed: one word per record (j below the 15 SIFt bits; 4 bytes up to 131072 atoms, 5 beyond),
    entity class recomputed on the host from the feat words: unpacks to exactly the sorted 16-byte stream."""
    p = arp_params.make_params()
    engine.set_params(p)
    soa = synth.cloud_featured(max(n, 1), seed=33)
    if n == 0:
        soa = AtomSoA(xyz=np.zeros((0, 3), np.float32), feat=np.zeros(0, np.uint32), res_id=np.zeros(0, np.int32),
                      rad_class=np.zeros(0, np.uint16), vdw=soa.vdw, cov=soa.cov, res_prev=np.zeros(0, np.int32),
                      res_next=np.zeros(0, np.int32), res_flags=np.zeros(0, np.uint8))
    engine.upload_atoms(soa)
    engine.run_pairs_async()
    pk = engine.fetch_pairs_packed(with_dist=True)
    exp = engine.fetch_pairs(pk.n, sorted=True)
    assert pk.n == exp.shape[0] and pk.n_faults == 0
    assert (pk.hi is not None) == (n > (1 << 17)) and pk.bits_j == max(1, int(np.ceil(np.log2(max(soa.n_atoms, 2)))))
    assert pk.nbytes == 4 * (soa.n_atoms + 1) + (9 if n > (1 << 17) else 8) * pk.n
    util.assert_records_equal(pk.to_records(soa.feat), exp, f'packed n={n}')
    pk2 = engine.fetch_pairs_packed(with_dist=False)
    util.assert_records_equal(pk2.to_records(soa.feat, dist=engine.fetch_pairs_dist(pk2.n)), exp, f'packed + distances on demand n={n}')


def test_packed_stream_counts_fault_records(engine):
    """The xbond-without-neighbour fault bit does not fit the packed word: the records that carry it are counted."""
    g = util.Golden('xbond_fault')
    engine.set_params(g.params)
    engine.upload_atoms(g.soa)
    engine.run_pairs_async()
    pk = engine.fetch_pairs_packed()
    full = engine.fetch_pairs(pk.n, sorted=True)
    n_fault = int(np.count_nonzero(full['mask'] & np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR)))
    assert n_fault > 0 and pk.n_faults == n_fault
    got = pk.to_records(g.soa.feat, dist=engine.fetch_pairs_dist(pk.n))
    full = full.copy()
    full['mask'] &= ~np.uint32(abi.PAIR_FAULT_XBOND_NO_NBR)
    util.assert_records_equal(got, full, 'packed stream of a run with fault records')


@pytest.mark.parametrize('case', ['plain', 'h_fix', 'pinned', 'no_optional', 'empty', 'big'])
def test_wire_forms_give_the_same_records(engine, case):
    """soa.WireAtoms (uint8 counts, sparse halogen neighbours, fixed-point hydrogens) decoded on the device: the records of
    the plain AtomSoA, bit for bit."""
    from arpeggio_b200.engine import pinned_soa
    p = arp_params.make_params()
    engine.set_params(p)
    if case == 'no_optional':
        soa = synth.cloud_uniform(4000, seed=3)
    elif case == 'empty':
        base = synth.cloud_featured(8, seed=1)
        z = np.zeros(0, np.int32)
        soa = AtomSoA(xyz=np.zeros((0, 3), np.float32), feat=np.zeros(0, np.uint32), res_id=z, rad_class=np.zeros(0, np.uint16),
                      vdw=base.vdw, cov=base.cov, res_prev=z, res_next=z, res_flags=np.zeros(0, np.uint8),
                      bond_off=np.zeros(1, np.int32), bond_nbr=z, h_off=np.zeros(1, np.int32), h_xyz=np.zeros((0, 3)),
                      xnbr_xyz=np.zeros((0, 3), np.float32))
    else:
        soa = synth.cloud_featured(70_000 if case == 'big' else 6000, seed=41, h_decimals=3 if case in ('h_fix', 'pinned', 'big') else None)
    w = soa.to_wire()
    assert (w.h_fix is not None) == (case in ('h_fix', 'pinned', 'big'))
    if case == 'pinned':
        w = pinned_soa(w)
    exp = oracle.pairs(soa, p)
    util.assert_records_equal(engine.pairs(w), exp, f'wire {case}')
    if case == 'pinned':
        assert engine.stats()['input_bytes'] == w.input_bytes()
    util.assert_records_equal(engine.pairs(soa), exp, f'plain after wire {case}')       # and back: the buffers change hands
    util.assert_records_equal(engine.pairs(w), exp, f'wire after plain {case}')


@pytest.mark.parametrize('case', [c for c in CASES if c != 'xbond_fault'])
def test_wire_forms_golden(engine, case):
    g = util.Golden(case)
    engine.set_params(g.params)
    w = g.soa.to_wire()
    rec = engine.pairs(w)
    util.assert_records_equal(rec, g.exp_pairs, f'wire {case}')


def test_wire_forms_are_checked(engine):
    import ctypes as C
    soa = synth.cloud_featured(500, seed=2)
    w = soa.to_wire()
    w.bond_cnt = w.bond_cnt.copy()
    w.bond_cnt[3] += 1                                  # no longer sums to n_bond_nbr
    with pytest.raises(Exception, match='bond_cnt'):
        engine.upload_atoms(w)
    w = soa.to_wire()
    if w.xnbr_idx.shape[0] >= 2:
        w.xnbr_idx = w.xnbr_idx[::-1].copy()
        with pytest.raises(Exception, match='xnbr_idx'):
            engine.upload_atoms(w)
    w = soa.to_wire()
    w.bond_off = soa.bond_off                           # both forms at once
    with pytest.raises(Exception, match='alternatives'):
        engine.upload_atoms(w)
    w = soa.to_wire()
    w.h_cnt = w.h_cnt.copy()
    w.h_cnt[-1] += 2                                    # more hydrogens announced than h_xyz holds
    with pytest.raises(Exception, match='h_cnt'):
        engine.upload_atoms(w)
    r = dataclasses.replace(soa, h_xyz=np.round(soa.h_xyz, 3)).to_wire()
    assert r.h_fix is not None
    r.h_fix_scale = 0.0
    with pytest.raises(Exception, match='h_fix_scale'):
        engine.upload_atoms(r)
    w = soa.to_wire()
    w.xnbr_idx = w.xnbr_idx.copy()
    if w.xnbr_idx.shape[0]:
        w.xnbr_idx[-1] = soa.n_atoms                    # out of range
        with pytest.raises(Exception, match='xnbr_idx'):
            engine.upload_atoms(w)
    engine.set_params(arp_params.make_params())
    util.assert_records_equal(engine.pairs(soa.to_wire()), oracle.pairs(soa, arp_params.make_params()), 'after the rejected uploads')


@pytest.mark.parametrize('n,expect', [(0, 0), (7, 0), (9000, 0), (9000, 1000), (9000, 10**9), (140_000, 500_000)])
def test_packed_stream_enqueued_behind_the_run(engine, n, expect):
    """arp_pairs_fetch_packed_async / _wait: the sorted packed view built with the record count read on the device, the
    first `expect` words copied blindly and the rest by the wait: the stream of the plain fetch."""
    from arpeggio_b200.engine import PackedPairs
    p = arp_params.make_params()
    engine.set_params(p)
    soa = synth.cloud_featured(max(n, 1), seed=55)
    if n == 0:
        z = np.zeros(0, np.int32)
        soa = AtomSoA(xyz=np.zeros((0, 3), np.float32), feat=np.zeros(0, np.uint32), res_id=z, rad_class=np.zeros(0, np.uint16),
                      vdw=soa.vdw, cov=soa.cov, res_prev=z, res_next=z, res_flags=np.zeros(0, np.uint8))
    exp = oracle.pairs(soa, p)
    cap = exp.shape[0] + 100
    for with_dist in (False, True):
        out = PackedPairs(np.zeros(soa.n_atoms + 2, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.float32))
        engine.upload_atoms(soa)
        engine.run_pairs_async()
        engine.fetch_pairs_packed_async(out, expect, with_dist)
        pk = engine.fetch_pairs_packed_wait()
        assert pk.n == exp.shape[0] and pk.n_faults == 0
        got = pk.to_records(soa.feat, dist=None if with_dist else engine.fetch_pairs_dist(pk.n))
        util.assert_records_equal(got, exp, f'async packed n={n} expect={expect} dist={with_dist}')
    # after a finished run the same call works without the blind path; a destination that is too small is fetched again
    engine.run_pairs()
    small = PackedPairs(np.zeros(soa.n_atoms + 2, np.uint32), np.zeros(max(exp.shape[0] // 2, 0), np.uint32),
                        np.zeros(max(exp.shape[0] // 2, 0), np.uint8), None)
    engine.fetch_pairs_packed_async(small, 0, False)
    pk = engine.fetch_pairs_packed_wait()
    util.assert_records_equal(pk.to_records(soa.feat, dist=engine.fetch_pairs_dist(pk.n)), exp, f'async packed after run n={n}')
    engine.upload_atoms(soa)
    engine.run_pairs_async()
    engine.fetch_pairs_packed_async(small, 0, False)                 # blind and too small
    pk = engine.fetch_pairs_packed_wait()
    util.assert_records_equal(pk.to_records(soa.feat, dist=engine.fetch_pairs_dist(pk.n)), exp, f'async packed, small destination n={n}')


def test_async_packed_fetch_survives_an_overflowing_run():
    """A fresh context sizes its record buffer from a guess; a structure far denser than the guess overflows the run that
    the blind view was built on: the wait repeats the run and fetches the plain way."""
    from arpeggio_b200.engine import ContactEngine, PackedPairs
    p = arp_params.make_params()
    rng = np.random.default_rng(3)
    base = synth.cloud_featured(6000, seed=9)
    dense = dataclasses.replace(base, xyz=(rng.random((6000, 3)) * 14.0).astype(np.float32), bond_off=None, bond_nbr=None)
    exp = oracle.pairs(dense, p)
    assert exp.shape[0] > 16 * 6000 + 4096 + 1000          # beyond the first guess of arp_pairs_run_async
    with ContactEngine(0, p) as eng:
        out = PackedPairs(np.zeros(6002, np.uint32), np.zeros(exp.shape[0] + 8, np.uint32), None, None)
        eng.upload_atoms(dense)
        eng.run_pairs_async()
        eng.fetch_pairs_packed_async(out, 0, False)
        pk = eng.fetch_pairs_packed_wait()
        util.assert_records_equal(pk.to_records(dense.feat, dist=eng.fetch_pairs_dist(pk.n)), exp, 'async packed after an overflow')


def test_radius_table_follows_the_upload(engine):
    """The K x K radius-sum table is kept across uploads whose vdw / cov tables are bitwise the same and rebuilt when a
    value, the number of classes or vdw_comp changes."""
    p = arp_params.make_params()
    engine.set_params(p)
    a = synth.cloud_featured(5000, seed=91)
    b = dataclasses.replace(a, vdw=a.vdw * 1.07)
    c = dataclasses.replace(a, cov=np.concatenate([a.cov * 0.9, [0.5]]), vdw=np.concatenate([a.vdw, [1.1]]))
    for name, soa in (('a', a), ('a again', a), ('b: other vdw radii', b), ('a', a), ('c: one class more', c), ('b', b)):
        util.assert_records_equal(engine.pairs(soa), oracle.pairs(soa, p), name)
    p2 = arp_params.make_params(5.0, 0.35, False)
    engine.set_params(p2)
    util.assert_records_equal(engine.pairs(b), oracle.pairs(b, p2), 'b after vdw_comp changed')
    assert not np.array_equal(oracle.pairs(a, p)['mask'], oracle.pairs(b, p)['mask'])
