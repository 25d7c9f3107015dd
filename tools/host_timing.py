#!/usr/bin/env python
"""Host time inside the entry points of the end-to-end step (ARPEGGIO_HOST_TIMING=1 makes arp_destroy print it) for the
pipelined packed stream of BatchRunner, by slots and submission threads.  Run under gpurun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['ARPEGGIO_HOST_TIMING'] = '1'
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.batch import BatchRunner  # noqa: E402
from arpeggio_b200.engine import pinned_soa  # noqa: E402

p = params.make_params()
soa = synth.cloud_featured(100_000, seed=2, h_decimals=3)
for name, host in (('plain', pinned_soa(soa)), ('wire', pinned_soa(soa.to_wire()))):
    for slots, threads in ((1, 1), (2, 1), (3, 1), (4, 1), (6, 1), (8, 1), (12, 1), (16, 1), (24, 1), (4, 2), (6, 2), (6, 3), (8, 2), (16, 2)):
        with BatchRunner(device=0, slots=slots, params=p, submit_threads=threads) as runner:
            runner.run([host] * 24, check_finite=False, packed=True)
            _, dt = runner.run([host] * 600, check_finite=False, packed=True)
            print(f'--- {name} slots={slots} submit_threads={threads}: {dt / 600 * 1e3:.4f} ms/step; per-engine host timing follows (slot 0 first)',
                  file=sys.stderr, flush=True)
