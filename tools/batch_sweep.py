#!/usr/bin/env python
"""configs[4] batch leg of bench.py by stream form, slots and submission threads.  Run under gpurun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.batch import BatchRunner  # noqa: E402
from arpeggio_b200.engine import pinned_soa  # noqa: E402

p = params.make_params()
n_struct, pack = 768, int(sys.argv[1]) if len(sys.argv) > 1 else 16
plain = [synth.cloud_featured(20_000, seed=1000 + k) for k in range(8)]
forms = {'plain': [pinned_soa(s) for s in plain], 'wire': [pinned_soa(s.to_wire()) for s in plain]}
for slots, threads in ((6, 1), (8, 1), (8, 2), (8, 4), (8, 8), (12, 4), (12, 6)):
    with BatchRunner(device=0, slots=slots, params=p, submit_threads=threads) as runner:
        for form, distinct in forms.items():
            shard = [distinct[k % 8] for k in range(n_struct)]
            for mode in ('packed', 'compact'):
                if mode == 'compact' and threads != 1:
                    continue                      # the compact path always runs one thread per slot
                kw = dict(packed=True) if mode == 'packed' else dict(compact=True)
                runner.run(shard[:slots * pack], check_finite=False, pack=pack, **kw)
                best = min(runner.run(shard, check_finite=False, pack=pack, **kw)[1] for _ in range(3))
                print(f'slots={slots:2d} threads={threads} {form:5s} {mode:7s} pack={pack}: {n_struct / best:9.0f} structures/s', flush=True)
