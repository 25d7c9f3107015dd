#!/usr/bin/env python
"""The pipelined end-to-end step (BatchRunner.run(packed=True), wire inputs, 8 streams) under the library's A/B knobs:
does a launch sequence that is fastest alone (cooperative grid kernel, k_classify spinning on k_search's hand-off, PDL) stay
the best one when eight streams share the device?  Run under gpurun: python tools/e2e_knobs.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %r)
from arpeggio_b200 import params, synth
from arpeggio_b200.batch import BatchRunner
from arpeggio_b200.engine import pinned_soa
p = params.make_params()
soa = synth.cloud_featured(100_000, seed=2)
host = pinned_soa(soa.to_wire())
with BatchRunner(device=0, slots=8, params=p) as runner:
    runner.run([host] * 24, check_finite=False, packed=True)
    best = min(runner.run([host] * 400, check_finite=False, packed=True)[1] for _ in range(3)) / 400
    e = runner.engines[0]
    e.upload_atoms(soa); e.run_pairs()
    print('%%.4f ms/step end to end   %%.4f ms device step alone' %% (best * 1e3, e.time_pairs(20, True)))
''' % ROOT

for knobs in ([], ['ARPEGGIO_NO_EARLY_CLASSIFY'], ['ARPEGGIO_NO_PDL'], ['ARPEGGIO_NO_REG_GRID'], ['ARPEGGIO_NO_FUSED_GRID'],
              ['ARPEGGIO_NO_EARLY_CLASSIFY', 'ARPEGGIO_NO_PDL'], ['ARPEGGIO_NO_EARLY_CLASSIFY', 'ARPEGGIO_NO_FUSED_GRID'],
              ['ARPEGGIO_NO_EARLY_CLASSIFY', 'ARPEGGIO_NO_PDL', 'ARPEGGIO_NO_FUSED_GRID'], ['ARPEGGIO_TILES']):
    env = dict(os.environ, **{k: '1' for k in knobs})
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    print(f'{"+".join(knobs) or "default":70s} {r.stdout.strip() or r.stderr[-300:]}', flush=True)
