#!/bin/bash
# tools/ab_env.sh "VAR=1 VAR2=x" ...: time the resident configs[2] job with the in-tree library under each environment
# setting (run under gpurun; A/B tooling, never a bench number).  ATOMS / STEPS override the defaults.
for e in "$@"; do
  env $e python - "$e" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
atoms = int(os.environ.get('ATOMS', 100000)); steps = int(os.environ.get('STEPS', 300))
soa = synth.cloud_featured(atoms, seed=2)
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa); n = eng.run_pairs()
    eng.time_pairs(20, flush_l2=True)
    ms = eng.time_pairs(steps, flush_l2=True); st = eng.stats()
    print('%-28s %7d atoms us/step %6.1f | grid %5.1f pairs %5.1f | search %5.1f classify %5.1f hscan %5.1f (kernels apart) | %.2f Gpairs/s' % (
        sys.argv[1], atoms, ms * 1e3, st['ms_grid'] * 1e3, st['ms_pairs'] * 1e3, st['ms_search'] * 1e3,
        (st['ms_classify'] - st['ms_hscan']) * 1e3, st['ms_hscan'] * 1e3, n / ms / 1e6), flush=True)
PY
done
