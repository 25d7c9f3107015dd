#!/usr/bin/env python
"""Three end-to-end steps on one stream (wire-form upload, pair kernels, sorted packed view, copies) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e2e.csv python tools/profile_e2e.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.batch import BatchRunner  # noqa: E402
from arpeggio_b200.engine import pinned_soa  # noqa: E402

atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
soa = synth.cloud_featured(atoms, seed=2, h_decimals=3)
host = pinned_soa(soa.to_wire())
with BatchRunner(device=0, slots=1, params=params.make_params()) as runner:
    counts, _ = runner.run([host] * 3, check_finite=False, packed=True)
    print(counts)
    parts = [pinned_soa(synth.cloud_featured(20_000, seed=1000 + k).to_wire()) for k in range(4)]
    counts, _ = runner.run(parts * 2, check_finite=False, packed=True, pack=4)
    print(counts)
