"""Batch runs on the GPU: several stream slots give exactly the per-structure results."""
import numpy as np
import pytest

import util
from arpeggio_b200 import params as arp_params, synth
from arpeggio_b200.batch import BatchRunner
from arpeggio_b200.soa import AtomSoA
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_batch_runner_matches_oracle(engine):
    p = arp_params.make_params()
    soas = [synth.cloud_featured(n, seed=50 + k) for k, n in enumerate((4000, 9000, 1500, 12000, 300, 7000, 2500))]
    got = {}
    with BatchRunner(device=engine.device, slots=3, params=p) as runner:
        counts, dt = runner.run(soas, consume=lambda i, rec: got.__setitem__(i, rec.copy()), sorted=True)
        assert dt > 0
        for i, soa in enumerate(soas):
            exp = oracle.pairs(soa, p)
            assert counts[i] == exp.shape[0]
            util.assert_records_equal(got[i], exp, f'structure {i}')
        # a second pass over the same contexts (buffers reused, different sizes per slot)
        counts2, _ = runner.run(soas[::-1])
        assert counts2 == counts[::-1]


def test_batch_runner_compact_stream(engine):
    """compact=True: every structure comes back as CompactPairs in the slot's pinned block (one wait per structure)
    and unpacks to the oracle's sorted stream; with and without the distance stream."""
    p = arp_params.make_params()
    soas = [synth.cloud_featured(n, seed=80 + k) for k, n in enumerate((300, 9000, 1500, 12000, 4000))]
    for with_dist in (True, False):
        got = {}
        with BatchRunner(device=engine.device, slots=2, params=p) as runner:
            counts, _ = runner.run(soas, consume=lambda i, cp: got.__setitem__(i, cp.to_records()), compact=True, with_dist=with_dist)
        for i, soa in enumerate(soas):
            exp = oracle.pairs(soa, p)
            assert counts[i] == exp.shape[0]
            if not with_dist:
                exp = exp.copy()
                exp['dist'] = 0
            util.assert_records_equal(got[i], exp, f'structure {i} with_dist={with_dist}')


def _local(cp):
    rec = cp.to_records()
    rec['i'] += 0                      # rows are local already (row_off is rebased); j is batch-global
    rec['j'] -= cp.atom_base
    return rec


def test_device_packed_batches(engine):
    """arp_upload_atoms_batch: structures with their own residue numbering, bond / hydrogen tables and radius tables
    (some sharing classes, some not, one without bonds, one without hydrogens, an empty one) are concatenated on the
    device; every structure's stream equals its stand-alone oracle stream."""
    p = arp_params.make_params()
    engine.set_params(p)
    soas = [synth.cloud_featured(n, seed=90 + k, bonds=(k != 2)) for k, n in enumerate((900, 4000, 1500, 1, 2500, 7000))]
    soas[3] = synth.cloud_featured(5, seed=99)
    # another radius table for one structure (reordered classes + a new one) and no hydrogens for another
    import dataclasses
    perm = np.array([5, 3, 0, 1, 2, 4])
    soas[1] = dataclasses.replace(soas[1], vdw=np.append(soas[1].vdw[perm], 2.3), cov=np.append(soas[1].cov[perm], 1.4),
                                  rad_class=np.argsort(perm)[soas[1].rad_class].astype(np.uint16))
    soas[4] = dataclasses.replace(soas[4], h_off=None, h_xyz=None)
    exp = [oracle.pairs(s, p) for s in soas]
    off = engine.upload_atoms_batch(soas)
    engine.run_pairs_async()
    cp = engine.fetch_pairs_compact(with_dist=True)
    assert cp.n == sum(e.shape[0] for e in exp)
    for k, e in enumerate(exp):
        util.assert_records_equal(_local(cp.structure(int(off[k]), int(off[k + 1]))), e, f'packed structure {k}')
    # the same through the batch runner, groups of 4 and of 3 (last group short), 2 slots
    got = {}
    with BatchRunner(device=engine.device, slots=2, params=p) as runner:
        for pack in (4, 3):
            got.clear()
            counts, _ = runner.run(soas, consume=lambda i, part: got.__setitem__(i, _local(part)), compact=True, with_dist=True, pack=pack)
            for k, e in enumerate(exp):
                assert counts[k] == e.shape[0]
                util.assert_records_equal(got[k], e, f'pack={pack} structure {k}')


def test_batch_runner_packed_stream(engine):
    p = arp_params.make_params()
    soas = [synth.cloud_featured(n, seed=85 + k) for k, n in enumerate((300, 9000, 1500, 12000))]
    got = {}
    with BatchRunner(device=engine.device, slots=2, params=p) as runner:
        counts, _ = runner.run(soas, consume=lambda i, pk: got.__setitem__(i, pk.to_records(soas[i].feat)), packed=True, with_dist=True)
    for i, soa in enumerate(soas):
        exp = oracle.pairs(soa, p)
        assert counts[i] == exp.shape[0]
        util.assert_records_equal(got[i], exp, f'packed structure {i}')


def test_device_packed_batch_of_wire_and_plain_structures(engine):
    """arp_upload_atoms_batch with structures in wire form, in plain form and mixed in one batch (counts win: the merged
    offsets come from one scan on the device)."""
    p = arp_params.make_params()
    engine.set_params(p)
    soas = [synth.cloud_featured(n, seed=300 + k, h_decimals=3 if k % 2 else None) for k, n in enumerate((900, 5, 2500, 0, 1200, 4097))]
    z = np.zeros(0, np.int32)
    soas[3] = AtomSoA(xyz=np.zeros((0, 3), np.float32), feat=np.zeros(0, np.uint32), res_id=z, rad_class=np.zeros(0, np.uint16),
                      vdw=soas[0].vdw, cov=soas[0].cov, res_prev=z, res_next=z, res_flags=np.zeros(0, np.uint8))
    exp = oracle.pairs(AtomSoA.concat(soas), p)
    for mix in ('wire', 'mixed'):
        parts = [s.to_wire() if (mix == 'wire' or k % 3 == 0) and s.n_atoms else s for k, s in enumerate(soas)]
        off = engine.upload_atoms_batch(parts)
        assert off[-1] == sum(s.n_atoms for s in soas)
        n = engine.run_pairs()
        util.assert_records_equal(engine.fetch_pairs(n, sorted=True), exp, f'batch {mix}')


@pytest.mark.parametrize('threads', [1, 2])
def test_batch_runner_packed_groups(engine, threads):
    """run(packed=True, pack=4): groups concatenated on the device, the packed stream of the batch cut into one view per
    structure whose records carry structure-local indices; wire-form inputs, more atoms per group than 131072 (fifth byte)."""
    p = arp_params.make_params()
    soas = [synth.cloud_featured(n, seed=500 + k, h_decimals=3) for k, n in enumerate((300, 9000, 0 + 1, 1500, 12000, 70_000, 70_000, 5, 2000))]
    got = {}
    with BatchRunner(device=engine.device, slots=3, params=p, submit_threads=threads) as runner:
        for wire in (False, True):
            src = [s.to_wire() for s in soas] if wire else soas
            got.clear()
            counts, _ = runner.run(src, consume=lambda i, pk: got.__setitem__(i, pk.to_records(soas[i].feat)), packed=True, with_dist=True, pack=4)
            for i, soa in enumerate(soas):
                exp = oracle.pairs(soa, p)
                assert counts[i] == exp.shape[0]
                util.assert_records_equal(got[i], exp, f'packed group member {i} wire={wire}')


def test_packed_words_of_a_batch_hold_structure_local_indices(engine):
    """A batch of structures none larger than 131072 atoms needs no fifth byte however many atoms it has in total; the
    whole batch unpacks with its atom offsets to the records of the concatenated structures."""
    p = arp_params.make_params()
    engine.set_params(p)
    soas = [synth.cloud_featured(n, seed=700 + k) for k, n in enumerate((60_000, 50_000, 45_000, 1))]
    off = engine.upload_atoms_batch(soas)
    engine.run_pairs_async()
    pk = engine.fetch_pairs_packed(with_dist=True)
    assert pk.bits_j == 16 and pk.hi is None and pk.n_atoms == 155_001
    whole = AtomSoA.concat(soas)
    exp = oracle.pairs(whole, p)
    util.assert_records_equal(pk.to_records(whole.feat, struct_off=off), exp, 'batch, global indices')
    # the same through one arp_atoms with struct_off (concatenated on the host)
    engine.upload_atoms(whole)
    engine.run_pairs_async()
    pk = engine.fetch_pairs_packed(with_dist=True)
    assert pk.bits_j == 16 and pk.hi is None
    util.assert_records_equal(pk.to_records(whole.feat, struct_off=whole.struct_off), exp, 'host-concatenated batch')
