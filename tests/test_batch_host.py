"""Host-side logic of batch runs without a GPU: sharding of the structure list over ranks, and the
world_size-2 reduction bench.py performs (gloo, CPU)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from arpeggio_b200 import synth
from arpeggio_b200.batch import shard_indices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_shards_partition_the_list(world):
    rng = np.random.default_rng(world)
    sizes = rng.integers(100, 50_000, size=257)
    parts = [shard_indices(sizes, world, r) for r in range(world)]
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(257))
    loads = [int(sizes[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= int(sizes.max())          # longest-first greedy bound
    assert parts == [shard_indices(sizes, world, r) for r in range(world)]   # deterministic


def test_shard_edge_cases():
    assert shard_indices([], 4, 0) == []
    assert shard_indices([10], 4, 0) == [0] and shard_indices([10], 4, 3) == []
    with pytest.raises(ValueError):
        shard_indices([1, 2], 2, 2)


def test_world_size_2_gloo_reduction(tmp_path):
    """Two CPU ranks shard a structure list, 'process' it, and reduce (max time, sum pairs) the way
    bench.py does under torchrun."""
    script = tmp_path / 'rank.py'
    script.write_text(textwrap.dedent(f'''
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from arpeggio_b200.batch import shard_indices
        rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        dist.init_process_group('gloo', rank=rank, world_size=world)
        sizes = np.arange(1, 41) * 100
        mine = shard_indices(sizes, world, rank)
        pairs = float(sum(13 * sizes[i] for i in mine))       # stand-in for the per-structure record counts
        secs = 1.0 + rank
        t = torch.tensor([secs], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([pairs, float(len(mine))], dtype=torch.float64); dist.all_reduce(s, op=dist.ReduceOp.SUM)
        if rank == 0:
            print('RESULT', float(t[0]), float(s[0]), int(s[1]))
        dist.barrier(); dist.destroy_process_group()
    '''))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29613')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29613', str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('RESULT')][0].split()
    assert float(line[1]) == 2.0
    assert float(line[2]) == 13 * 100 * sum(range(1, 41))
    assert int(line[3]) == 40


def test_make_selection_can_stay_with_the_host_class():
    """cuda_make_selection = False: the mixin hands _make_selection to the class it is mixed into (the reference's own
    list(set(...)) construction, whose order depends on its KD-tree pair order) and forgets its packed image."""
    from arpeggio_b200.dropin import CudaContactsMixin

    class Base:
        def _make_selection(self, selections):
            self.seen = list(selections)

    class Host(CudaContactsMixin, Base):
        cuda_make_selection = False

    h = Host()
    h._cuda_pack_cache = ('stale', None)
    h._make_selection(['/A/508/'])
    assert h.seen == ['/A/508/'] and h._cuda_pack_cache is None and h._cuda_pack_version == 1


def test_wire_atoms_are_a_lossless_image_of_the_soa():
    """soa.WireAtoms: uint8 counts for the two offset arrays, sparse halogen neighbours, int32 fixed-point hydrogens only
    when every coordinate survives the round trip; the ctypes image carries the totals the library checks."""
    from arpeggio_b200 import abi
    from arpeggio_b200.soa import WireAtoms
    soa = synth.cloud_featured(3000, seed=12)
    w = soa.to_wire()
    assert isinstance(w, WireAtoms) and w.n_atoms == soa.n_atoms and w.bond_off is None and w.h_off is None
    assert np.array_equal(np.concatenate([[0], np.cumsum(w.bond_cnt, dtype=np.int64)]), soa.bond_off)
    assert np.array_equal(np.concatenate([[0], np.cumsum(w.h_cnt, dtype=np.int64)]), soa.h_off)
    has = (soa.feat & np.uint32(abi.F_HAS_XNBR)) != 0
    assert np.array_equal(w.xnbr_idx, np.flatnonzero(has)) and np.array_equal(w.xnbr_xyz, soa.xnbr_xyz[has])
    if w.h_fix is not None:
        assert w.h_xyz is None and np.array_equal(w.h_fix / w.h_fix_scale, soa.h_xyz)
    else:
        assert w.h_xyz is soa.h_xyz and w.h_fix_scale == 0.0
    # hydrogens that are not 3-decimal fractions stay float64; 3-decimal ones go to fixed point
    import dataclasses
    off = dataclasses.replace(soa, h_xyz=soa.h_xyz + 1e-7)
    assert off.to_wire().h_fix is None and off.to_wire().h_xyz is not None
    rounded = dataclasses.replace(soa, h_xyz=np.round(soa.h_xyz, 3))
    wr = rounded.to_wire()
    assert wr.h_fix is not None and wr.h_fix.dtype == np.int32 and np.array_equal(wr.h_fix / 1000.0, rounded.h_xyz)
    assert wr.input_bytes() < 0.6 * rounded.input_bytes()
    ct = wr.as_ctypes()
    assert ct.n_bond_nbr == soa.bond_nbr.shape[0] and ct.n_h == soa.h_xyz.shape[0] and ct.n_xnbr == int(has.sum())
    assert ct.h_fix_scale == 1000.0 and not ct.bond_off and not ct.h_off and not ct.h_xyz and ct.bond_cnt and ct.h_cnt and ct.h_fix
