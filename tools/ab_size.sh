#!/bin/bash
# tools/ab_size.sh [label]: time the resident configs[2] job with one library (run under gpurun; A/B tooling, never
# a bench number).  LIB = library to load (default: the in-tree build), ATOMS / STEPS override the defaults.
ARPEGGIO_CUDA_LIB=${LIB:-$PWD/arpeggio_b200/libarpeggio_cuda.so} python - "${1:-cur}" <<'PY'
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
    print('%-8s %7d atoms us/step %6.1f | grid %5.1f pairs %5.1f | search %5.1f classify %5.1f hscan %5.1f (kernels apart) | %.2f Gpairs/s' % (
        sys.argv[1], atoms, ms * 1e3, st['ms_grid'] * 1e3, st['ms_pairs'] * 1e3, st['ms_search'] * 1e3,
        (st['ms_classify'] - st['ms_hscan']) * 1e3, st['ms_hscan'] * 1e3, n / ms / 1e6), flush=True)
PY
