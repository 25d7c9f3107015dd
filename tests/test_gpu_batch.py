"""Batch runs on the GPU: several stream slots give exactly the per-structure results."""
import numpy as np
import pytest

import util
from arpeggio_b200 import params as arp_params, synth
from arpeggio_b200.batch import BatchRunner
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
